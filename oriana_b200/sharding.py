"""Row (cell) sharding across the GPUs of one box.

Rank r owns a contiguous block of cells: X, a1, a2, U_hat.  The gene side (b1, b2, V_hat, pi, alpha, beta)
is replicated and recomputed redundantly on every rank from all-reduced sums, so the only traffic per CAVI
iteration is two sum-allreduces: [Zj | D^T U_hat] (2 p K float32) and [colsum D_hat | sum_i log U_hat |
sum_i U_hat | ELBO partials] (p + 2K + 8 float64).  `torch.distributed` (NCCL on GPUs; gloo in the CPU
tests of this host logic) is plumbing only.
"""
import torch
import torch.distributed as dist


class RowSharding:

    def __init__(self, group=None, enabled=False):
        self.group = group
        if enabled:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError('sharded=True needs an initialised torch.distributed process group')
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.enabled = self.world > 1

    def allreduce_sum(self, tensor):
        """In-place sum over ranks (stream-ordered on the current CUDA stream for NCCL)."""
        if self.enabled:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def broadcast(self, tensor, src=0):
        """Make a replicated (gene-side) tensor identical on every rank: rank `src`'s values win."""
        if self.enabled:
            dist.broadcast(tensor, src=src, group=self.group)
        return tensor

    def total_rows(self, n_rows, device=None):
        if not self.enabled:
            return int(n_rows)
        t = torch.tensor([int(n_rows)], dtype=torch.int64, device=device or 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    @staticmethod
    def row_block(n_total, rank, world):
        """Contiguous [begin, end) of cells owned by `rank`; block sizes differ by at most one."""
        base, extra = divmod(int(n_total), int(world))
        begin = rank * base + min(rank, extra)
        return begin, begin + base + (1 if rank < extra else 0)
