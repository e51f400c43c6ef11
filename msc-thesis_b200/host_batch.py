"""Host-resident batches: stream chunks of episodes through the GPU with copies and launches overlapped.

A batch that lives in (pinned) host memory -- because it is larger than the device, or because the caller owns
it -- is cut into chunks; `n_slots` device-side ClosedLoop3DOF instances, each on its own CUDA stream, take
turns: host->device copy of the chunk's state, one fused launch, device->host copy of the results.  Copy engines
and SMs work concurrently, so the end-to-end rate approaches max(copy time, compute time) instead of their sum."""
import torch


def loop_state_tensors(loop):
    """Every mutable per-episode tensor of a ClosedLoop3DOF, in a fixed order (what a host batch consists of)."""
    env, rls = loop.env, loop.rls
    return [env.state, env.t_ep, loop.ref2, env.corr_u, loop.pid.hdot_corr, rls._W, rls._cov, rls.x, rls.u,
            rls._skip, loop.agent_col.net, loop.agent_lon.net, env.step_count, env.flags, env.term_step, loop.acc]


def loop_result_tensors(loop):
    """Per-episode results of a launch: accumulated returns, flags, termination tick, tick counter + statistics."""
    return [loop.acc, loop.env.flags, loop.env.term_step, loop.env.step_count, loop.stats_dev]


I_COV, I_LON = 6, 11            # positions of rls._cov and agent_lon.net in loop_state_tensors


def _triu_rows(k=9):
    """Rows of the (k*k, N) covariance block that hold its upper triangle, and for every row of the full block the
    position of its value in that list."""
    up = [i * k + j for i in range(k) for j in range(i, k)]
    pos = {r: p for p, r in enumerate(up)}
    full = [pos[min(i, j) * k + max(i, j)] for i in range(k) for j in range(k)]
    return up, full


class HostBatchPipeline:
    def __init__(self, make_loop, chunk, n_slots=2, compact=False):
        """make_loop(n) -> ClosedLoop3DOF over n episodes on the current device.

        compact: the host copy of a chunk omits what is redundant -- the lower triangle of the covariance while it is
        symmetric (36 of 81 rows) and the target critic of the cyclic agent while tau = 1 makes it equal to the critic
        (24 of 66 rows): 17 % fewer bytes over PCIe; both are re-created on the device after the upload.  Measured on one
        B200 (1,048,576 episodes, 8 chunks): SLOWER, 80.6 ms per pass against 61.6 ms -- the small re-creation kernels of
        one slot queue behind the other slot's fused kernel, which owns every SM's registers, and the upload / launch
        overlap is lost; so it is off by default (it only pays where the host link, not the GPU, is the limit)."""
        self.chunk = int(chunk)
        self.slots = [make_loop(self.chunk) for _ in range(n_slots)]
        self.streams = [torch.cuda.Stream() for _ in range(n_slots)]
        self.t0 = self.slots[0].env.t
        self.k0 = self.slots[0].k
        lp = self.slots[0]
        tau = lp.agent_lon.tau_target_critic
        self.pack_cov = bool(compact and lp.rls.assume_symmetric)
        self.pack_lon = bool(compact and not torch.is_tensor(tau) and float(tau) == 1.0)
        dev = lp.device
        up, full = _triu_rows()
        self._up = torch.tensor(up, device=dev)
        self._full = torch.tensor(full, device=dev)
        nh = lp.agent_lon.n_hidden
        self._lon_rows = 7 * nh
        # device staging buffers for the compact parts, one set per slot
        self._stage = [(torch.empty(len(up), self.chunk, dtype=s.rls._cov.dtype, device=dev) if self.pack_cov else None,
                        torch.empty(7 * nh, self.chunk, dtype=s.agent_lon.net.dtype, device=dev) if self.pack_lon else None)
                       for s in self.slots]

    def _host_view(self, i, t):
        """The part of state tensor i that travels."""
        if i == I_COV and self.pack_cov:
            return t[self._up.to(t.device)]
        if i == I_LON and self.pack_lon:
            return t[:self._lon_rows]
        return t

    def host_chunks_from(self, loop0, n_chunks):
        """Pinned host copies (one list of tensors per chunk) of a loop's current state, replicated."""
        src = [self._host_view(i, t) for i, t in enumerate(loop_state_tensors(loop0))]
        return [[torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in src] for _ in range(n_chunks)]

    def result_buffers(self, n_chunks):
        res = loop_result_tensors(self.slots[0])
        return [[torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in res] for _ in range(n_chunks)]

    def run(self, host_chunks, host_results, n_ticks):
        """One pass over the host batch: every chunk is uploaded, advanced by n_ticks ticks and its results are
        downloaded.  Returns after all work has been ENQUEUED; synchronise before reading host_results."""
        main = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(main)
        for c, (src, dst) in enumerate(zip(host_chunks, host_results)):
            slot = self.slots[c % len(self.slots)]
            stream = self.streams[c % len(self.streams)]
            stage_cov, stage_lon = self._stage[c % len(self.slots)]
            with torch.cuda.stream(stream):
                for i, (h, d) in enumerate(zip(src, loop_state_tensors(slot))):
                    if i == I_COV and self.pack_cov:
                        stage_cov.copy_(h, non_blocking=True)
                        torch.index_select(stage_cov, 0, self._full, out=d)         # mirror the upper triangle
                    elif i == I_LON and self.pack_lon:
                        stage_lon.copy_(h, non_blocking=True)
                        nh = self._lon_rows // 7
                        d[:7 * nh].copy_(stage_lon)
                        d[7 * nh:].copy_(stage_lon[3 * nh:])                        # target critic == critic (tau = 1)
                    else:
                        d.copy_(h, non_blocking=True)
                slot.k, slot.env.t = self.k0, self.t0
                slot.run(n_ticks)
                for h, d in zip(dst, loop_result_tensors(slot)):
                    h.copy_(d, non_blocking=True)
        for s in self.streams:
            main.wait_stream(s)
