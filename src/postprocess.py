from gnomix_b200.postprocess import get_meta_data, write_msp, write_fb  # noqa: F401  (reference: src/postprocess.py:25-126)
