from gnomix_b200.base import Base  # noqa: F401  (reference: src/Base/base.py:8)
