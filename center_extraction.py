"""`import center_extraction as ce` shim, as in the reference's python-prototype/ (main.py:23):
re-exports the B200 implementation in upmix_b200/center_extraction.py."""
from upmix_b200.center_extraction import *  # noqa: F401,F403
from upmix_b200.center_extraction import (EPS, MultiBandExtractorAccu, chain_bands,  # noqa: F401
                                          extract_center_left_right_multi_band_in_memory, main)

if __name__ == "__main__":
    main()
