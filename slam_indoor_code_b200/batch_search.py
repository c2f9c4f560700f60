"""Host mirror of the reference's search for the next good frame in a batch
(cycleProcessing/batch.cpp:101-226), on top of the one-call batch matcher.

The reference matches the previous frame against the batch elements one by one, last element first
(`findGoodFramesFromBatchSingleThread`, :120-148), or speculatively on `threadsCount` threads
(`...MultiThreads`, :162-226, then `findGoodFrameFromMatchedBatch`, :270-316); both apply the same
selection rule.  Here every element is matched in ONE call (Context.matchBatch ->
slamb200_match_batch), then that rule picks the frame.  No CPU fallback for the matching.
"""
FRAME_NOT_FOUND = -1   # batch.h:6


def selectGoodFrameFromMatchCounts(matched, requiredMatchedPointsCount, useFirstFitInBatch,
                                   skipFramesFromBatchHead=0):
    """batch.cpp:120-148 / :283-307: walk from the last element down to skipFramesFromBatchHead; an
    element is good when it has >= requiredMatchedPointsCount matches and >= the matches of the good
    element so far; first-fit stops at the first good one.  The reference compares a size_t with an
    int: a negative requirement becomes a huge unsigned number and nothing is good."""
    good, good_size = FRAME_NOT_FOUND, 0
    required = int(requiredMatchedPointsCount)
    if required < 0:
        required += 1 << 64
    for i in range(len(matched) - 1, max(int(skipFramesFromBatchHead), 0) - 1, -1):
        m = int(matched[i])
        if m >= required and m >= good_size:
            good, good_size = i, m
            if useFirstFitInBatch:
                break
    return good


def findGoodFrameFromBatch(ctx, previousDescriptor, batchDescriptors, matcherType,
                           requiredMatchedPointsCount, useFirstFitInBatch, skipFramesFromBatchHead=0,
                           knnMatcherDistance=0.7):
    """Returns (goodIndex or FRAME_NOT_FOUND, allMatches): allMatches[i] is what the reference stores
    in BatchElement::matches of element i; the caller takes allMatches[goodIndex] as `matches` and
    drops the batch up to goodIndex (batch.cpp:91-98)."""
    all_matches = ctx.matchBatch(previousDescriptor, batchDescriptors, matcherType, knnMatcherDistance)
    good = selectGoodFrameFromMatchCounts([len(m) for m in all_matches], requiredMatchedPointsCount,
                                          useFirstFitInBatch, skipFramesFromBatchHead)
    return good, all_matches
