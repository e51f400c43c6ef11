"""Which NCCL debug settings make the communicator's INFO lines visible? (diagnostic for bench.py's N > 1 path)"""
import os
import sys
mode = sys.argv[1] if len(sys.argv) > 1 else "stderr"
print("rank", os.environ.get("RANK"), "NCCL_DEBUG at entry:", os.environ.get("NCCL_DEBUG"), "FILE:", os.environ.get("NCCL_DEBUG_FILE"), flush=True)
if mode != "asis":
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
    if mode == "stderr":
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(8, device="cuda")
dist.all_reduce(x)
torch.cuda.synchronize()
print("rank", dist.get_rank(), "sum", float(x[0]), flush=True)
dist.destroy_process_group()
