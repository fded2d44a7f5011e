from gnomix_b200.smooth import host_slide_window as _hsw


def slide_window(B, S, y=None):
    """reference: src/Smooth/utils.py:4-29 (float32 [N*W, S*A] rows, labels flattened)."""
    N, W, A = B.shape
    return _hsw(B, S), (None if y is None else y.reshape(N * W))
