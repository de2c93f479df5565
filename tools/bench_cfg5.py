"""BASELINE config 5: ONE 4 s CPI @ 20 MS/s (N = 8e7, 512 delay x 4097 Doppler) split over the GPUs of a box, clutter
filter (410 taps) included -- blah2_b200.shard.SingleCpiPlan over the C-ABI communicator (NCCL from C).

    python tools/bench_cfg5.py [iters] [--no-filter]                                   (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cfg5.py [iters]

Each rank holds only its 1/N slice of the IQ.  Prints one JSON line (rank 0): ms per CPI (CUDA events on the compute
stream, max over ranks), and -- with --check -- the map's error against the numpy oracle on a SMALLER geometry."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from blah2_b200.process import Ambiguity, WienerHopfChunk
from blah2_b200.shard import Comm, SingleCpiPlan

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    iters = int(args[0]) if args else 10
    use_filter = "--no-filter" not in sys.argv
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm(rank, world, local)
    geom = (0, 511, -512, 512, 20000000, 80000000, True)
    clutter = (-10, 400) if use_filter else None
    amb = Ambiguity(*geom, device=local)
    g = amb.geometry
    plan = SingleCpiPlan(comm, amb, geom[5], torch.device("cuda", local), clutter=clutter,
                         whc_factory=lambda a, b, n, c0, nc: WienerHopfChunk(a, b, n, c0, nc, device=local))
    # synthetic scene slice: reference = noise, surveillance = 0.5 x + delayed copies + noise (so the filter has work)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.randn((plan.ns, 2), device="cuda", generator=gen)) * 1000.0
    y = 0.5 * x + 0.1 * torch.roll(x, 3) + 0.05 * torch.roll(x, 17) + torch.view_as_complex(torch.randn((plan.ns, 2), device="cuda", generator=gen)) * 10.0
    plan.x_own.copy_(x)
    plan.y_own.copy_(y)
    del x, y
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            m = plan.run(s)
    comm.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(iters):
            m = plan.run(s)
        if comm.rank != 0:
            comm.join(s)
        e1.record(s)
    comm.sync()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # per-stage times of one more CPI (rank 0's stream; exchanges include waiting for the slowest rank)
    marks = []
    with torch.cuda.stream(s):
        plan.run(s, marks)
    comm.sync()
    torch.cuda.synchronize()
    stages = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 4) for i in range(1, len(marks))}
    ok = plan.whc.last_status() if plan.whc is not None else True
    if rank == 0:
        byts = (2 * 16 * geom[5] if use_filter else 0) + 16 * g.n_used + 8 * g.n_doppler_bins * g.n_delay_bins
        print(json.dumps({"config": "cfg5 single 4 s CPI @ 20 MS/s, 512 x 4097" + (", WienerHopf 410 taps" if use_filter else ", CAF only"),
                          "n_gpus": world, "ms_per_cpi": round(float(ms), 4), "maps_per_s": round(1e3 / float(ms), 2),
                          "msamples_per_s": round(geom[5] / float(ms) / 1e3, 1),
                          "algorithmic_GBps_aggregate": round(byts / float(ms) / 1e6, 1), "filter_ok": bool(ok),
                          "range_fft": g.range_fft_len, "segments": g.range_segments, "doppler_fft": g.doppler_fft_len,
                          "stage_ms_rank0": stages, "map_finite": bool(torch.isfinite(torch.view_as_real(m)).all()), "map_absmax": float(m.abs().max())}), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
