from .unsorted_segment_sum import UnsortedSegmentSumFunction, unsorted_segment_sum  # noqa: F401
