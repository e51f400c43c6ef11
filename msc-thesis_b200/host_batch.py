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


class HostBatchPipeline:
    def __init__(self, make_loop, chunk, n_slots=2):
        """make_loop(n) -> ClosedLoop3DOF over n episodes on the current device."""
        self.chunk = int(chunk)
        self.slots = [make_loop(self.chunk) for _ in range(n_slots)]
        self.streams = [torch.cuda.Stream() for _ in range(n_slots)]
        self.t0 = self.slots[0].env.t
        self.k0 = self.slots[0].k

    def host_chunks_from(self, loop0, n_chunks):
        """Pinned host copies (one list of tensors per chunk) of a loop's current state, replicated."""
        src = loop_state_tensors(loop0)
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
            with torch.cuda.stream(stream):
                for h, d in zip(src, loop_state_tensors(slot)):
                    d.copy_(h, non_blocking=True)
                slot.k, slot.env.t = self.k0, self.t0
                slot.run(n_ticks)
                for h, d in zip(dst, loop_result_tensors(slot)):
                    h.copy_(d, non_blocking=True)
        for s in self.streams:
            main.wait_stream(s)
