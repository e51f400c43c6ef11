"""Functional + cost probe of TMEM used as thread-private storage (include/heli3dof.h: h3_tmem_probe)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msc_thesis_b200 as h3  # noqa: E402
from msc_thesis_b200 import _lib as L  # noqa: E402


def run(warps, blocks, reps):
    dev = torch.device("cuda", 0)
    nt = blocks * warps * 32
    out = torch.zeros(137, nt, dtype=torch.float64, device=dev)
    cyc = torch.zeros(2, blocks, dtype=torch.int64, device=dev)
    # cycles are stored [2*block + i]
    cyc_flat = cyc.view(-1)
    L.check(L.lib().h3_tmem_probe(warps, blocks, reps, L.ptr(out), L.ptr(cyc_flat), L.stream_ptr(dev)))
    torch.cuda.synchronize()
    gt = torch.arange(nt, dtype=torch.float64, device=dev)
    want = gt[None, :] * 1024 + torch.arange(128, dtype=torch.float64, device=dev)[:, None]
    ok_fill = bool(torch.equal(out[:128], want))
    ok_rmw = bool(torch.equal(out[128:136], want[:8] + reps))
    c = cyc_flat.view(blocks, 2).double()
    return {"warps": warps, "blocks": blocks, "reps": reps, "fill_readback_exact": ok_fill, "rmw_exact": ok_rmw,
            "cycles_per_rmw_pass_120_words": float(c[:, 0].mean() / max(reps, 1)),
            "cycles_per_dependent_load": float(c[:, 1].mean() / max(reps * 8, 1))}


if __name__ == "__main__":
    res = [run(4, 1, 64), run(8, 1, 64), run(4, 148 * 2, 64), run(8, 148, 64), run(4, 148 * 8, 16)]
    print(json.dumps(res, indent=1))
