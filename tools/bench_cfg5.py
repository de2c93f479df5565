"""BASELINE config 5: ONE 4 s CPI @ 20 MS/s (N = 8e7, 512 delay x 4097 Doppler) split over the GPUs of a box.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cfg5.py [iters]
Each rank holds only its 1/N slice of the IQ; one all-gather of the range matrix; column-tiled Doppler
stage; NCCL gather of the map tiles to rank 0.  Prints one JSON line (rank 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from blah2_b200.process import Ambiguity
from blah2_b200.shard import block_range, caf_single_cpi_sharded

def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    geom = (0, 511, -512, 512, 20000000, 80000000, True)
    amb = Ambiguity(*geom, device=local)
    g = amb.geometry
    b0, nb = block_range(g.n_doppler_bins, rank, world)
    n_local = nb * g.n_corr
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.randn((n_local, 2), device="cuda", generator=gen))
    y = torch.view_as_complex(torch.randn((n_local, 2), device="cuda", generator=gen))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            m = caf_single_cpi_sharded(amb, x, y, rank, world, s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(iters):
            m = caf_single_cpi_sharded(amb, x, y, rank, world, s)
        e1.record(s)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        byts = 16 * g.n_used + 8 * g.n_doppler_bins * g.n_delay_bins
        print(json.dumps({"config": "cfg5 single 4 s CPI @ 20 MS/s, 512 x 4097", "n_gpus": world, "ms_per_cpi": round(float(ms), 4),
                          "maps_per_s": round(1e3 / float(ms), 2), "msamples_per_s": round(geom[5] / float(ms) / 1e3, 1),
                          "algorithmic_GBps_aggregate": round(byts / float(ms) / 1e6, 1),
                          "range_fft": g.range_fft_len, "segments": g.range_segments, "doppler_fft": g.doppler_fft_len,
                          "map_shape": list(m.shape)}), flush=True)
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
