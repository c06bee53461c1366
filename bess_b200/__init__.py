"""bess_b200: B200-native (sm_100a CUDA) implementation of the BeSS primal-dual active-set hot path behind the
reference's own entry points (pywrap_bess / bessCpp).  See DESIGN.md and include/bess_b200.h."""
__all__ = ["cbess", "linear", "engine", "gen_data", "build", "compat", "dist"]
