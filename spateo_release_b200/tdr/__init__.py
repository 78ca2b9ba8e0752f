"""``st.tdr`` entry points that share the Gaussian-kernel vector field (reference: spateo/tdr/__init__.py)."""

from .morphofield import morphofield, morphofield_gp, morphofield_sparsevfc
