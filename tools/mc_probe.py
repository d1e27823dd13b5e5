"""probe: does torch symmetric memory expose an NVSwitch multicast mapping on this box?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty(1 << 20, dtype=torch.uint8, device="cuda")
h = symm_mem.rendezvous(t, dist.group.WORLD)
print("rank", rank, "world", h.world_size, "multicast_ptr", hex(getattr(h, "multicast_ptr", 0) or 0), "has_multicast_support", getattr(h, "has_multicast_support", None), flush=True)
dist.barrier(); dist.destroy_process_group()
