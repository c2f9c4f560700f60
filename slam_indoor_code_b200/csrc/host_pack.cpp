// host_pack.cpp -- lossless narrowing of a SIFT descriptor Mat on the host before it crosses PCIe.
//
// cv::SIFT emits CV_32F rows whose values are integers in [0, 255], so a 5.12 MB
// Mat of 10 000 rows carries 1.28 MB of information.  slamb200_upload_desc_packed narrows the rows
// to bytes on the calling thread, straight into page-locked staging, verifying EVERY element
// (integer, in range) and every row norm (< 2^20, the exact-mode condition of the tcgen05 path);
// the GPU then reads a quarter of the bytes.  A Mat that fails the check is uploaded as fp32.
// AVX2 when the CPU has it, otherwise a scalar loop.  Compiled by the host compiler (not nvcc).
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#if defined(__x86_64__)
#include <immintrin.h>

template <bool NT>
__attribute__((target("avx2"))) static int pack_rows_avx2(const float* src, size_t stride, int n,
                                                          uint8_t* dst) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  const __m256i lim = _mm256_set1_epi32(255);
  int norm_bad = 0;
  for (int r = 0; r < n; r++) {
    const float* s = src + (size_t)r * stride;
    uint8_t* d = dst + (size_t)r * 128;
    __m256i sq = _mm256_setzero_si256();
    for (int i = 0; i < 128; i += 32) {
      const __m256 f0 = _mm256_loadu_ps(s + i), f1 = _mm256_loadu_ps(s + i + 8),
                   f2 = _mm256_loadu_ps(s + i + 16), f3 = _mm256_loadu_ps(s + i + 24);
      const __m256i i0 = _mm256_cvttps_epi32(f0), i1 = _mm256_cvttps_epi32(f1),
                    i2 = _mm256_cvttps_epi32(f2), i3 = _mm256_cvttps_epi32(f3);
      // not an integer (or NaN / out of int range: cvtt gives 0x80000000, caught by the range test)
      const __m256 ne = _mm256_or_ps(
          _mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(i0), f0, _CMP_NEQ_UQ),
                       _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), f1, _CMP_NEQ_UQ)),
          _mm256_or_ps(_mm256_cmp_ps(_mm256_cvtepi32_ps(i2), f2, _CMP_NEQ_UQ),
                       _mm256_cmp_ps(_mm256_cvtepi32_ps(i3), f3, _CMP_NEQ_UQ)));
      const __m256i any = _mm256_or_si256(_mm256_or_si256(i0, i1), _mm256_or_si256(i2, i3));
      bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_castps_si256(ne), _mm256_andnot_si256(lim, any)));
      const __m256i p01 = _mm256_packs_epi32(i0, i1), p23 = _mm256_packs_epi32(i2, i3);
      sq = _mm256_add_epi32(sq, _mm256_add_epi32(_mm256_madd_epi16(p01, p01), _mm256_madd_epi16(p23, p23)));
      const __m256i p = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(p01, p23), perm);
      // NT: the staging buffer is written once and next read by the GPU over PCIe -- streaming
      // stores spare the read-for-ownership of every destination line (a quarter of the bytes
      // this loop moves) and keep the fp32 source from being evicted by its own output
      if (NT) _mm256_stream_si256((__m256i*)(d + i), p);
      else _mm256_storeu_si256((__m256i*)(d + i), p);
    }
    __m128i h = _mm_add_epi32(_mm256_castsi256_si128(sq), _mm256_extracti128_si256(sq, 1));
    h = _mm_add_epi32(h, _mm_shuffle_epi32(h, 0x4e));
    h = _mm_add_epi32(h, _mm_shuffle_epi32(h, 0xb1));
    norm_bad |= _mm_cvtsi128_si32(h) >= (1 << 20);
    // a Mat of general floats is recognised within its first rows: stop reading it
    if ((r & 63) == 63 && (norm_bad || !_mm256_testz_si256(bad, bad))) return 0;
  }
  if (NT) _mm_sfence();
  return _mm256_testz_si256(bad, bad) && !norm_bad;
}
#endif

static int pack_rows_scalar(const float* src, size_t stride, int n, uint8_t* dst) {
  int ok = 1;
  for (int r = 0; r < n; r++) {
    const float* s = src + (size_t)r * stride;
    uint8_t* d = dst + (size_t)r * 128;
    int ss = 0;
    for (int i = 0; i < 128; i++) {
      const float f = s[i];
      const int v = (f >= 0.f && f <= 255.f) ? (int)f : -1;
      if (v < 0 || (float)v != f) { ok = 0; d[i] = 0; continue; }
      d[i] = (uint8_t)v;
      ss += v * v;
    }
    if (ss >= (1 << 20)) ok = 0;
  }
  return ok;
}

// rows: n x 128 floats, `stride` floats apart; dst: n x 128 bytes.  Returns 1 when every element is
// an integer in [0, 255] and every squared row norm is below 2^20 (dst is then the exact image of
// the rows), 0 otherwise (dst is unspecified).
extern "C" int slamb200_host_pack_u8(const float* src, size_t stride, int n, uint8_t* dst) {
#if defined(__x86_64__)
  static const int have_avx2 = __builtin_cpu_supports("avx2");
  static const int nt = [] { const char* e = getenv("SLAMB200_PACK_NT"); return e ? atoi(e) : 1; }();
  if (have_avx2) {
    if (nt && ((uintptr_t)dst & 31) == 0) return pack_rows_avx2<true>(src, stride, n, dst);
    return pack_rows_avx2<false>(src, stride, n, dst);
  }
#endif
  return pack_rows_scalar(src, stride, n, dst);
}
