from .cmatrix import CountMatrix
from .generation import generate_factor_matrices, generate_factor_matrices_device, generate_u, generate_v
from .synth import synth_counts_device

__all__ = ['CountMatrix', 'generate_factor_matrices', 'generate_factor_matrices_device', 'generate_u', 'generate_v', 'synth_counts_device']
