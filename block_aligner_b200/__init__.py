"""block_aligner_b200 -- B200-native batched adaptive-block affine-gap aligner.

Host-side mirror of the reference's interface over the C ABI of libblock_aligner_b200.so
(see include/block_aligner_b200.h). Import is cheap; the shared library is loaded on first use.
"""
from .api import (LOCAL_START, FREE_QUERY_START_GAPS,  # noqa: F401
                  Aligner, AAProfile, AlignResult, BaConfig, BaStats, Batch, Block, BlockAlignerError, Cigar, Gaps,  # noqa: F401
                  Library, PaddedBytes, SizeRange, SCORING_AA, SCORING_BYTE, SCORING_NUC, SCORING_PROFILE, TRACE,
                  XDROP, LOCAL_START, FREE_QUERY_START_GAPS, aa_matrix_simple, concat, nuc_matrix, runs_to_string)
