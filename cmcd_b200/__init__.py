"""cmcd_b200 -- B200-native (sm_100a) implementation of the CMCD bridge hot path.

Host-side mirror of the reference's operator surface for that path
(mcdboundingmachine / boundingmachine / mcd_utils / model_handler / nn / variationaldist /
utils.log_final_losses), calling hand-written CUDA through the C ABI in
include/cmcd_b200.h (libcmcd_b200.so).  There is no CPU or eager fallback: every compute
entry point raises if the CUDA library or a CUDA device is missing.
"""
__version__ = "0.1.0"
