"""SparseGaP (oriana/models/sparse_gap.py) -- present so that the reference's drivers import cleanly
(`experiments/clustering.py:5` imports it next to the models it uses).  In the reference its `step()` cannot run: the
E-step reads an undefined name (sparse_gap.py:127, NameError), so there is no behaviour to reproduce; constructing it
here says so instead of failing later.  The sparse model that works in the reference is `SparseZIGaP`."""


class SparseGaP:

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('SparseGaP.step() raises NameError in the reference (sparse_gap.py:127): there is no '
                                  'reference behaviour to reproduce; use SparseZIGaP (sparsity + dropout) or GaP')
