"""Probe torch symmetric memory on this box (run under torchrun, 2+ GPUs)."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        print("handle attrs:", [a for a in dir(hdl) if not a.startswith("_")])
        print("buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "dev array", hex(hdl.buffer_ptrs_dev), "signal pads", [hex(p) for p in hdl.signal_pad_ptrs])
    t.fill_(rank + 1.0)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    print(rank, "peer value", float(peer[0]), "has multicast", getattr(hdl, "has_multicast_support", None), "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0))
    hdl.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
dist.destroy_process_group()
