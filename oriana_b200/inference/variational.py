"""Mean-field variational family bookkeeping (oriana/inference/variational.py:6-23)."""


class VariationalDistribution:
    """List of (model node, variational node) pairs; one pair per mean-field factor."""

    def __init__(self):
        self._partitions = []

    def add_partition(self, node_p, node_q):
        node_q.name = node_p.name + '-variational'
        pair = (node_p, node_q)
        if pair not in self._partitions:
            self._partitions.append(pair)

    @property
    def partitions(self):
        return list(self._partitions)

    def __len__(self):
        return len(self._partitions)
