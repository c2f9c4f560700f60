# GPU parity tests under every developer switch of the library (tail forms incl. the experimental in-kernel
# tail as the default, PCIe-read uploads without streaming stores, general-float sub-batches of two)
T="tests/test_gpu_matching.py tests/test_gpu_full_parity.py tests/test_gpu_api_contract.py tests/test_gpu_tail_forms.py tests/test_gpu_host_cpp.py"
for cfg in "SLAMB200_TAIL_FORM=3" "SLAMB200_TAIL_FORM=2" "SLAMB200_UPLOAD_DMA=0 SLAMB200_PACK_NT=0" "SLAMB200_GEN_SUB=2 SLAMB200_TAIL_FORM=1"; do
echo "== $cfg"; env $cfg timeout 900 python -m pytest $T -m gpu -x -q 2>&1 | tail -2
done
