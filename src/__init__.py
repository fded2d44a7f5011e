"""Import-path compatibility with AI-sandbox/gnomix: the reference's driver (`gnomix.py`) and its
whole-object pickles refer to `src.model.Gnomix`, `src.Base.models.*`, `src.Smooth.models.*`,
`src.utils`, `src.postprocess` (gnomix.py:11-21, 206-209).  These modules re-export the
gnomix_b200 classes under those paths; nothing is implemented here."""
