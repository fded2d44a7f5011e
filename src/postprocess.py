from gnomix_b200.postprocess import get_meta_data, write_msp, write_fb  # noqa: F401  (reference: src/postprocess.py:25-126)
from gnomix_b200.postprocess import msp_to_lai, get_bed_data, msp_to_bed  # noqa: F401  (reference: src/postprocess.py:128-210)
