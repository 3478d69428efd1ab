"""pylabfea_b200 -- B200-native (sm_100a) implementation of pyLabFEA's per-element elastic-plastic stress update
(Material.response and its callees) and global stiffness assembly (Model.setupK), behind the reference's own
Model / Material API.  Host code is Python; all numerics run in hand-written CUDA kernels reached through the
C-ABI of libplfb200.so (include/plfb200.h) via ctypes.  No PyTorch, no Triton, no CPU fallback."""
from .basic import yf_tolerance, sig_dev, eps_eq, sig_eq_j2
from .material import Material
from .model import Model

__version__ = '0.1.0'
__all__ = ['Material', 'Model', 'yf_tolerance', 'sig_dev', 'eps_eq', 'sig_eq_j2']
