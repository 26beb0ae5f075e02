"""Copy-only probe of the host <-> device path under N ranks (torchrun): every rank copies the
benchmark's per-step traffic (435 MB of pinned records up, 1.07 GB image down) back to back,
nothing else running.  The aggregate over ranks is the ceiling the e2e leg can reach on this
box; compare with `e2e.h2d_bytes_per_step + d2h_bytes_per_step` per step.

    python -m torch.distributed.run --nproc-per-node N profiles/pcie_probe.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '.')
from katsdpimager_b200 import _lib, accel      # noqa: E402


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo', rank=rank, world_size=world)
    context = accel.Context(local)
    queue = context.create_command_queue()
    up, down = 435456000, 1073741824
    host_up = accel.HostArray((up,), np.uint8, context=context)
    host_down = accel.HostArray((down,), np.uint8, context=context)
    host_up[:] = 1
    host_down[:] = 0
    dev_up = accel.DeviceArray(context, (up,), np.uint8)
    dev_down = accel.DeviceArray(context, (down,), np.uint8)
    reps = 10

    def loop():
        for _ in range(reps):
            _lib.call('kib_memcpy_h2d_async', dev_up.ptr, host_up.ctypes.data, up, queue.stream)
            _lib.call('kib_memcpy_d2h_async', host_down.ctypes.data, dev_down.ptr, down, queue.stream)
    loop()
    queue.finish()
    if dist is not None:
        dist.barrier()
    start = queue.enqueue_marker()
    loop()
    stop = queue.enqueue_marker()
    stop.wait()
    seconds = stop.time_since(start)
    if dist is not None:
        import torch
        t = torch.tensor([seconds], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t[0])
    if rank == 0:
        per_rank = reps * (up + down) / seconds / 1e9
        print(json.dumps({'ranks': world, 'ms_per_step_traffic': seconds / reps * 1e3,
                          'gb_per_s_per_rank': per_rank, 'gb_per_s_aggregate': per_rank * world,
                          'bytes_per_step': up + down}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
