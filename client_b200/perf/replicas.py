"""Multi-GPU = independent replicas: one load generator per GPU,
own regions / streams / connections, no data-path collective.  The only cross-rank
traffic is measurement plumbing: a barrier around the timed region, the max over ranks
of the elapsed time and the sum of completed requests.  ``torch.distributed`` carries
it (``nccl`` on GPUs, ``gloo`` in the CPU tests)."""

import os


class Replicas:
    def __init__(self, backend=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._dist = None
        self._device = None
        if self.world > 1:
            import torch
            import torch.distributed as dist

            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self._device = "cuda:%d" % self.local_rank
            else:
                self._device = "cpu"
            if not dist.is_initialized():
                dist.init_process_group(backend)
            self._dist = dist

    # -- partitioning ---------------------------------------------------------------------
    def stream_base(self, slots_per_rank):
        """First Philox stream id of this rank: ranks own disjoint id ranges so no two
        GPUs ever generate the same tensor."""
        return self.rank * (1 << 40) + 0 * slots_per_rank

    def seed(self, base_seed):
        return base_seed + 1000003 * self.rank

    # -- measurement plumbing -----------------------------------------------------------------
    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()
            if self._device and self._device.startswith("cuda"):
                import torch

                torch.cuda.synchronize(self.local_rank)

    def _reduce(self, value, op):
        if self._dist is None:
            return value
        import torch

        t = torch.tensor([float(value)], dtype=torch.float64, device=self._device)
        self._dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, value):
        return self._reduce(value, self._dist.ReduceOp.MAX if self._dist else None)

    def sum(self, value):
        return self._reduce(value, self._dist.ReduceOp.SUM if self._dist else None)

    def throughput(self, local_units, local_seconds):
        """Whole-job throughput: all units of all ranks / the slowest rank's time."""
        return self.sum(local_units) / self.max(local_seconds)

    def walk(self, plan, enabled=True):
        """Run ``plan`` -- [(name, fn(sync))] -- on every rank in step: each entry meets the other ranks at
        exactly two barriers whatever happens on this rank, the one ``fn`` calls through ``sync`` (between "my
        instance is warm" and "time now"; made up for here when ``fn`` fails before it, or is skipped) and one
        after the entry.  After the first failure the remaining entries are skipped, the barriers are not.
        Returns ({name: result}, error string or None)."""
        results, error = {}, None
        for name, fn in plan:
            met = [False]

            def sync_once():
                if not met[0]:
                    met[0] = True
                    self.barrier()

            if enabled and error is None:
                try:
                    results[name] = fn(sync_once)
                except Exception as ex:
                    error = "%s: %s: %s" % (name, type(ex).__name__, ex)
            sync_once()
            self.barrier()
        return results, error

    def close(self):
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
