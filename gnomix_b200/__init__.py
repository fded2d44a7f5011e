"""gnomix_b200 -- B200-native (sm_100a) local-ancestry inference hot path behind the
Gnomix model API (AI-sandbox/gnomix src/model.py): Base.predict_proba ->
Smoother.predict_proba/predict -> gnofix.  Python host over the C ABI in
include/gnx.h (libgnx.so, hand-written CUDA).  No CPU fallback."""
from .gbt import GBTForest  # noqa: F401
from .base import Base, LogisticRegressionBase, CovRSKBase  # noqa: F401
from .smooth import Smoother, XGB_Smoother, CRF_Smoother  # noqa: F401
from .model import Gnomix  # noqa: F401
from .pickle_compat import load_model  # noqa: F401

__version__ = "0.1.0"
