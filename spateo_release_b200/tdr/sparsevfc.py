"""SparseVFC solver behind ``st.tdr.morphofield_sparsevfc`` (device implementation lands with the vfc_sweep kernel)."""


def morphofield_sparsevfc_core(*args, **kwargs):
    raise NotImplementedError("the SparseVFC device solver is not implemented yet in spateo_release_b200")
