"""Probe: does torch symmetric memory (and NVLS multicast) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; world", hdl.world_size, "multicast_ptr", hex(hdl.multicast_ptr),
          "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:3], "signal pads",
          len(hdl.signal_pad_ptrs), "pad size", hdl.signal_pad_size, flush=True)
except Exception as exc:  # noqa: BLE001
    print(rank, "symmetric memory failed:", repr(exc)[:400], flush=True)
dist.destroy_process_group()
