from gnomix_b200.model import Gnomix  # noqa: F401  (reference: src/model.py:12)
