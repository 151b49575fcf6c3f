from .cmatrix import CountMatrix
from .synth import synth_counts_device

__all__ = ['CountMatrix', 'synth_counts_device']
