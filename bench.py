#!/usr/bin/env python
"""bench.py — samples/s of the path-tracing core on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--workload c2_full|c1|c3_full|c4_stress]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] = C2, the "dragon-class" 869 880-triangle mesh instanced as Disney
metal and glass over a textured ground quad under an importance-sampled 2048x1024 HDR environment, 1280x720, max depth 6
(scenes/gen_scenes.py: c2_full; synthetic, generated on the spot and flattened by the reference's unchanged loader + BVH
builder).  One STEP = one pass of the hot path over one batch: --spp-per-step (32) samples for every pixel of the frame
(29.5 M pixel-samples); the default 32 steps are the config's full 1024 spp.

  value      whole-job samples/s, scene resident in HBM, timed with CUDA events on the launching stream, barrier +
             synchronize on both sides, max over ranks.  N > 1: every rank renders its own frame numbers (frame = RNG seed,
             disjoint streams) of every step, weak scaling; the accumulation buffers are summed once with NCCL inside the
             timed region.
  e2e        the same metric through the C ABI with HOST buffers: every step uploads its uniforms (LfParams + LfCamera,
             host structs), renders, and reads the accumulated image back to host memory (W*H*3 floats).
  roofline   extend kernel (closest-hit traversal, the dominant kernel): algorithmic bytes (SURVEY.md 8d formula with the
             kernel's own visit counts) / its mean launch duration from CUDA events inside the timed region, against the
             measured HBM copy peak of MEASURED_PEAKS.json; the L2 figure north_star asks for is reported beside it.
  cpu_baseline   the CPU oracle port (oracle/lf_oracle.cpp, OpenMP, all host cores) on a bounded sample of the same workload.
  --impl reference   the reference's path on the host CPU: the oracle port on this workload (C2 exceeds the buffer-texture
             limit of the only runnable GL, llvmpipe), plus the UNMODIFIED reference renderer on llvmpipe timed on C1.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (generator, description, spp of the full config)
    "c2_full": ("c2_full", "C2 dragon-class 869880-tri mesh x2 (Disney metal + glass) + textured ground + 2048x1024 HDR env, 1280x720, depth 6", 1024),
    "c1": ("cornell_256", "C1 cornell_box.scene 256x256, depth 4", 64),
    "c3_full": ("c3_full", "C3 instanced multi-mesh, textured materials, 2 quad + 1 sphere light, 1920x1080, depth 4", 256),
    "c4_stress": ("c4_stress", "C4 1296 x glass_sphere.obj = 20.57M instanced triangles, depth 8, 3840x2160", 64),
}
METRIC = "samples/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_pack(name):
    from scenes import gen_scenes
    out = os.path.join(ROOT, "scenes", "_gen", name)
    pack = os.path.join(out, f"{name}.lfpack")
    if not os.path.exists(pack):
        log(f"[bench] generating scene {name} ...")
        t0 = time.time()
        gen_scenes.build_pack(name, out)
        log(f"[bench] scene ready in {time.time() - t0:.1f} s")
    return pack


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU side
def stratified_tiles(W, H, tile=(80, 45), grid=4):
    """A stratified sample of the frame for the CPU arms: the frame is cut into tiles of `tile` pixels (shrunk to divide the frame)
    and grid x grid of them are taken, one from the middle of every cell of a grid x grid partition of the tile lattice - sky, horizon,
    mesh and ground are all represented in proportion, unlike a single centre tile (which is the most expensive part of the frame)."""
    tw, th = min(tile[0], W), min(tile[1], H)
    while W % tw:
        tw -= 1
    while H % th:
        th -= 1
    ntx, nty = W // tw, H // th
    gx, gy = min(grid, ntx), min(grid, nty)
    tiles = sorted({(int((i + 0.5) * ntx / gx), int((j + 0.5) * nty / gy)) for i in range(gx) for j in range(gy)})
    return tw, th, tiles


class OracleSampler:
    """The oracle port on a bounded, stratified sample of the workload: `spp` frames of every sampled tile per call."""

    def __init__(self, pack_path):
        import oracle_api
        from oracle_api import Oracle
        # torchrun exports OMP_NUM_THREADS=1 to its ranks: ask for every host core explicitly and report what OpenMP really uses
        self.threads = oracle_api.set_threads(os.cpu_count() or 1)
        self.o = Oracle(pack_path)
        self.W, self.H = self.o.params.width, self.o.params.height
        self.tw, self.th, self.tiles = stratified_tiles(self.W, self.H)
        self.o.update_params(tile_width=self.tw, tile_height=self.th)
        self.frame = 2

    def render(self, spp):
        t0 = time.time()
        for tx, ty in self.tiles:
            self.o.render_frames(self.frame, spp, 1, tx, ty)
        self.frame += spp
        return time.time() - t0

    def samples(self, spp):
        return self.tw * self.th * len(self.tiles) * spp

    def describe(self, spp):
        frac = self.tw * self.th * len(self.tiles) / (self.W * self.H)
        return (f"oracle port (CPU restatement of the GLSL, OpenMP, {self.threads} threads): {len(self.tiles)} tiles of {self.tw}x{self.th} stratified "
                f"over the {self.W}x{self.H} frame ({frac:.3f} of its pixels), {spp} spp per step")

    def close(self):
        self.o.close()


def oracle_sample_rate(pack_path, budget_s):
    """cpu_baseline of the CUDA arm: the oracle port on the stratified sample, as many spp as fit the time budget."""
    sm = OracleSampler(pack_path)
    sm.render(1)                                       # first touch of the scene arrays
    t1 = sm.render(1)
    spp = int(max(2, min(512, budget_s / max(t1, 1e-3))))
    dt = sm.render(spp)
    sample = sm.describe(spp) + f" ({sm.samples(spp)} pixel-samples, {dt:.1f} s)"
    threads = sm.threads
    sm.close()
    return sm.samples(spp) / dt, sample, threads


def run_reference(args, workload, desc):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only; the other ranks exit without work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    pack = ensure_pack(WORKLOADS[workload][0])
    sm = OracleSampler(pack)
    sm.render(1)                                       # first touch of the scene arrays
    t1 = sm.render(1)
    # size one step so that warmup + steps finish within ~2.5 minutes
    per_step_budget = min(2.0, 150.0 / max(1, args.steps + args.warmup))
    spp = int(max(1, min(256, per_step_budget / max(t1, 1e-3))))
    for _ in range(args.warmup):
        sm.render(spp)
    dt = sum(sm.render(spp) for _ in range(args.steps))
    value = sm.samples(spp) * args.steps / dt
    sample = sm.describe(spp)
    W, H, threads = sm.W, sm.H, sm.threads
    sm.close()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "resolution": [W, H], "spp_per_step": spp, "sample": sample, "same_config": True,
                   "same_config_note": "same scene, resolution, depth and RNG frames as the CUDA arm; a stratified subset of its pixels (every host thread busy), rate in samples/s"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_llvmpipe:
        lp = llvmpipe_rate(workload) if workload in LLVMPIPE_SCENES else None
        if lp and "samples_per_s_steady" in lp:
            # the reference itself runs this workload's scene: it is the figure of this line; the oracle port's stays beside it
            line["oracle_port"] = dict(line["cpu_baseline"])
            v = lp["samples_per_s_steady"]
            what = f"the unmodified reference renderer + GLSL on Mesa llvmpipe ({lp['gl']}), steady state after the shader JIT: {lp['workload']}"
            line["value"] = v
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": lp["threads"], "host_cores": os.cpu_count(), "kind": "reference", "sample": what}
            line["e2e"]["value"] = v
            line["config"]["sample"] = what
            line["llvmpipe"] = lp
        else:
            line["llvmpipe_c1"] = llvmpipe_c1()
    print(json.dumps(line), flush=True)


# The reference on llvmpipe per workload: (scene generator, spp, what it is).  C3 / C4 are timed on the SAME scene at 480x270 (the
# cost of a pixel-sample does not depend on the frame size to first order; a full 4K frame costs llvmpipe minutes per sample);
# C2's 0.87 M-triangle mesh exceeds this llvmpipe's 65 536-texel buffer textures, so C2 has no llvmpipe figure (oracle port instead).
LLVMPIPE_SCENES = {
    "c1": ("cornell_256", 64, "C1 cornell_box.scene 256x256, depth 4, 64 spp"),
    "c3_full": ("c3_small", 6, "the C3 scene at 480x270 (same instances, textures, lights), 6 spp"),
    "c4_stress": ("c4_mini", 5, "the C4 scene (1296 instances, 20.57 M triangles, depth 8) at 480x270, 5 spp"),
}


def llvmpipe_rate(workload):
    """The UNMODIFIED reference renderer + GLSL on Mesa llvmpipe on this box's host cores: steady-state samples/s (all draws after
    the first, which pays the one-time shader JIT)."""
    from scenes import gen_scenes
    ref = os.path.join(ROOT, "oracle", "_ref", "lf_ref_llvmpipe")
    if workload not in LLVMPIPE_SCENES:
        return {"unavailable": f"{workload} does not fit this llvmpipe's texture-buffer limits"}
    if not os.path.exists(ref) or gen_scenes.mesa_dir() is None:
        return {"unavailable": "oracle/_ref/lf_ref_llvmpipe or the bundled Mesa libGL not present"}
    gen, spp, what = LLVMPIPE_SCENES[workload]
    try:
        scene = gen_scenes.SCENES[gen](os.path.join(ROOT, "scenes", "_gen", gen))
        res = subprocess.run([ref, "--scene", scene, "--spp", str(spp), "--out", f"/tmp/lf_{workload}_llvmpipe.f32", "--timing-json"],
                             env=gen_scenes.llvmpipe_env(), capture_output=True, text=True, timeout=900, check=True)
        j = json.loads(res.stdout.strip().splitlines()[-1])
        lpt = str(j["lp_num_threads"])
        threads = int(lpt) if lpt.isdigit() else min(os.cpu_count() or 1, 16)     # Mesa 18.1 llvmpipe: one rasteriser thread per core, at most LP_MAX_THREADS = 16
        return {"kind": "reference", "workload": what, "samples_per_s_steady": j["samples_per_s_steady"], "first_step_s": j["first_step_s"],
                "render_s": j["render_s"], "cores": os.cpu_count(), "threads": threads, "lp_num_threads": j["lp_num_threads"], "gl": j["gl_renderer"] + " / " + j["gl_version"]}
    except Exception as e:  # the reference arm must not take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def llvmpipe_c1():
    return llvmpipe_rate("c1")


class _DevBuf:
    """Zero-copy view of a device buffer owned by liblfcuda (for torch.distributed collectives)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


# ------------------------------------------------------------------------------------------------ CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2_full", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--kernel-mode", type=int, default=0, help="0 wavefront (default), 1 megakernel")
    ap.add_argument("--frames-in-flight", type=int, default=0, help="pixel-sample frames per wavefront batch (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-llvmpipe", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the c5_strong record (4K scene, 4096 spp in total, strong scaling)")
    ap.add_argument("--c5-spp", type=int, default=4096)
    args = ap.parse_args()
    gen_name, desc, full_spp = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, args.workload, desc)
        return

    import torch
    import torch.distributed as dist
    import lavaframe_b200 as lf
    from lavaframe_b200.pathtracer import algorithmic_bytes_split, algorithmic_bytes_total
    from lavaframe_b200.multigpu import rank_frames

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        log(f"[bench] WORLD_SIZE {world} != --gpus {args.gpus}; using WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    if rank == 0:
        pack_path = ensure_pack(gen_name)
    barrier()
    pack_path = ensure_pack(gen_name)
    pack = lf.ScenePack(pack_path)

    pt = lf.PathTracer(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    pt.set_stream(stream.cuda_stream)
    pt.upload_pack(pack, kernel_mode=args.kernel_mode, frames_in_flight=args.frames_in_flight)
    W, H = pt.params.width, pt.params.height
    S, K, Wm = args.spp_per_step, args.steps, args.warmup
    npix = W * H
    if world > 1:   # NCCL communicator of the library itself (lfcuda_reduce), id broadcast through torch.distributed
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pt.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        pt.nccl_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    def step_frames(i):
        """first frame / stride of this rank in step i: global frame numbers 2, 3, ... dealt round-robin to the ranks."""
        f0, n, st = rank_frames(2 + i * S * world, S * world, rank, world)
        assert n == S
        return f0, st

    def sync():
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up
    for i in range(Wm):
        f0, st = step_frames(i)
        pt.render_frames(f0, S, st)
    if world > 1:
        pt.reduce()                                    # warm-up of the collective too (NCCL connects lazily on first use)
    sync()

    # ---- timed region: K steps (+ one NCCL sum), device resident
    pt.clear()
    pt.set_profiling(True)
    pt.stage_stats()                                   # drop the warm-up's stage events
    stats0 = pt.stage_stats()
    launches0 = pt.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()                                    # before the barrier: forking nvidia-smi takes tens of ms, different on every
    barrier(); sync()                                  # rank, and would skew the ranks' start times inside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        f0, st = step_frames(Wm + i)
        pt.render_frames(f0, S, st)
    if world > 1:
        pt.reduce()
    e1.record(stream)
    sync(); barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = pt.launch_count() - launches0
    stats1 = pt.stage_stats()
    pt.set_profiling(False)
    ms = max_over_ranks(ms)
    if world > 1:
        tl = torch.tensor([launches], device="cuda", dtype=torch.int64); dist.all_reduce(tl); launches = int(tl.item())
    total_samples = npix * S * K * world
    value = total_samples / (ms * 1e-3)

    # ---- e2e through the C ABI with host buffers: uniforms in every step; the finished image out.  Every step uploads its uniforms
    # (host structs), renders, and reads the accumulated image back to host memory; at N > 1 the read-out of a step is the NCCL sum
    # (lfcuda_reduce, in place) of the ranks' buffers, so each rank's accumulator is cleared per step and the host adds the steps up -
    # the progressive image a viewer of the multi-GPU job would be shown after every step.
    host_img = torch.empty((H, W, 3), dtype=torch.float32).pin_memory().numpy()
    host_sum = np.zeros((H, W, 3), np.float32) if world > 1 else None
    h2d = ctypes.sizeof(lf.LfParams) + ctypes.sizeof(lf.LfCamera)
    d2h = npix * 3 * 4
    pt.clear()
    barrier(); sync()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(K):
        f0, st = step_frames(Wm + K + i)
        pt.set_params(pt.params); pt.set_camera(pack.camera())       # host structs -> device uniforms
        if world > 1:
            pt.clear()
        pt.render_frames(f0, S, st)
        if world > 1:
            pt.reduce()                                               # lfcuda_reduce: ncclAllReduce of the accumulation buffers
        pt.read_accum(host_img)                                       # D2H + synchronize (every rank holds the sum)
        if world > 1 and rank == 0:
            host_sum += host_img
    e1.record(stream)
    sync(); barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_value = total_samples / (e2e_ms * 1e-3)

    # ---- N > 1: proof that the NCCL sum is the image.  A bounded job (the first `cf` frames) is rendered split over the ranks and summed
    # by lfcuda_reduce, then rendered whole by rank 0 alone; the two must agree up to fp32 summation order.
    reduce_check = None
    if world > 1:
        cf = 8 * world
        f0, n, st = rank_frames(2, cf, rank, world)
        pt.clear(); pt.render_frames(f0, n, st); pt.reduce()
        reduced = pt.read_accum().copy()
        same_everywhere = torch.tensor([float(np.float64(reduced.sum(dtype=np.float64)))], device="cuda", dtype=torch.float64)
        lo, hi = same_everywhere.clone(), same_everywhere.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            pt.clear(); pt.render_frames(2, cf, 1)
            single = pt.read_accum()
            err = np.abs(reduced - single)
            tol = 2e-5 * np.abs(single) + 1e-5
            reduce_check = {"frames": cf, "ranks": world, "max_abs_err": float(err.max()), "max_rel_err": float((err / np.maximum(np.abs(single), 1e-3)).max()),
                            "allclose_rtol_2e-5": bool((err <= tol).all()), "all_ranks_hold_the_same_sum": bool(lo.item() == hi.item()),
                            "what": "lfcuda_reduce(sum of the ranks' strided frames) vs rank 0 rendering the same frames alone"}
            assert reduce_check["allclose_rtol_2e-5"] and reduce_check["all_ranks_hold_the_same_sum"], f"NCCL-reduced image != single-GPU image: {reduce_check}"
        barrier()

    # ---- roofline of the dominant kernel (extend): visit counts from an instrumented pass over the first timed steps
    ncount = min(K, 2)
    pt.update_params(count_work=1)
    pt.reset_counters(); pt.clear()
    for i in range(ncount):
        f0, st = step_frames(Wm + i)
        pt.render_frames(f0, S, st)
    c = pt.counters()
    pt.update_params(count_work=0)
    scale = K / ncount
    ext_bytes, sh_bytes = algorithmic_bytes_split(c)
    stage = {k: {"launches": stats1[k]["launches"] - stats0[k]["launches"], "ms": stats1[k]["ms"] - stats0[k]["ms"]} for k in stats1}
    dom = "megakernel" if args.kernel_mode == 1 else "extend"
    dom_bytes = (algorithmic_bytes_total(c) if args.kernel_mode == 1 else ext_bytes) * scale
    dom_ms, dom_launches = stage[dom]["ms"], max(1, stage[dom]["launches"])
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic = None
    prof = os.path.join(ROOT, "profiles", "extend_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    rays = c["rays_closest"] + c["rays_shadow"]
    # The walk is served by L1 / L2, not HBM (DRAM traffic = `traffic`, a few per cent of the algorithmic bytes), so the roofline that
    # bounds it is the L2 -> SM read rate, measured in this run (16 MiB working set, the 16-byte read-only loads the traversal uses).
    # `frac_of_hbm` (against MEASURED_PEAKS.json) stays beside it, and so does the ceiling of the bound ncu names (L1 data-pipe
    # wavefronts: one-node-per-lane 2 x 256-bit fetches that hit L2), also measured in this run.
    l2 = hbm_read = node_l2 = node_l1 = None
    try:
        l2 = pt.measure_read_bandwidth(16 << 20, 10)
        hbm_read = pt.measure_read_bandwidth(2 << 30, 3)
        node_l2 = pt.measure_node_fetch(8 << 20, 5)
        node_l1 = pt.measure_node_fetch(32 << 10, 5)
    except Exception as e:
        log(f"[bench] bandwidth probes failed: {e}")
    roofline = {
        "bound": "l2", "kernel": dom, "achieved": achieved, "peak": l2, "unit": "GB/s", "frac": achieved / l2 if l2 else None, "traffic": traffic,
        "peak_source": "L2 -> SM read rate measured in this run (lfcuda_measure_read_bandwidth, 16 MiB working set)",
        "frac_of_l2": achieved / l2 if l2 else None,
        "hbm_peak": hbm_peak, "hbm_peak_source": hbm_src, "frac_of_hbm": achieved / hbm_peak,
        "l2_read_gbs_measured": l2, "hbm_read_gbs_measured": hbm_read,
        "l1_node_fetch_ceiling_gbs": node_l2, "l1_node_fetch_l1hit_gbs": node_l1,
        "frac_of_l1_node_fetch_ceiling": achieved / node_l2 if node_l2 else None,
        "l1_node_fetch_source": "lfcuda_measure_node_fetch in this run: 64-byte records, one per lane, 2 x LDG.256, 8 MiB table (L2 hits) / 32 KiB table (L1 hits)",
        "launches": dom_launches, "ms_per_launch": dom_ms / dom_launches, "bytes_per_launch": dom_bytes / dom_launches,
        "bytes_per_ray": (ext_bytes + sh_bytes) / max(1, rays),
        "stage_ms": {k: round(v["ms"], 3) for k, v in stage.items()},
        "shadow_achieved_gbs": sh_bytes * scale / (stage["shadow"]["ms"] * 1e-3) / 1e9 if stage["shadow"]["ms"] > 0 else None,
    }
    mrays = rays * scale * world / (ms * 1e-3) / 1e6

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "resolution": [W, H], "spp_per_step": S, "spp_total": S * K * world, "full_config_spp": full_spp,
                   "max_depth": pt.params.max_depth, "parallelism": f"spp-split x{world}" if world > 1 else "single GPU",
                   "kernels": "megakernel" if args.kernel_mode == 1 else "wavefront",
                   "l2_policy": f"the path state one step touches ({S * W * H * 352 / 1e9:.1f} GB: spp_per_step x pixels x 352 B) and the scene geometry exceed the 126 MB L2 every step; no flush needed"},
        "mrays_per_s": mrays, "rays_per_sample": rays / max(1, c["samples"]),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K,
                "path": ("lfcuda_set_params + lfcuda_set_camera + lfcuda_render_frames + lfcuda_read_accum (host buffers)" if world == 1 else
                         "lfcuda_set_params + lfcuda_set_camera + lfcuda_clear + lfcuda_render_frames + lfcuda_reduce (NCCL) + lfcuda_read_accum (host buffers), every step, every rank")},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if reduce_check is not None:
        line["reduce_check"] = reduce_check
    pt.close()

    # ---- north_star's scaling config (C5): the 4K scene at 4096 spp in total, split over the ranks (STRONG scaling), NCCL sum inside
    if not args.no_c5 and args.workload == "c2_full" and args.kernel_mode == 0:
        try:
            line["c5_strong"] = c5_strong(args, rank, local_rank, world, stream, barrier, max_over_ranks)
        except Exception as e:   # the headline line must survive
            line["c5_strong"] = {"failed": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, sample, threads = oracle_sample_rate(pack_path, 12.0)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port", "sample": sample}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5_strong(args, rank, local_rank, world, stream, barrier, max_over_ranks):
    """BASELINE.json configs[4]: the 4K synthetic scene (C4) at 4096 spp IN TOTAL, the frames dealt round-robin to the ranks (4096 / N
    per rank), accumulation buffers summed by lfcuda_reduce (NCCL) inside the timed region.  Total work is fixed as N grows: the
    driver's N = 1, 2, 4, 8 runs of this record are north_star's strong-scaling figure (>= 7x at 8 GPUs)."""
    import torch
    import torch.distributed as dist
    import lavaframe_b200 as lf
    from lavaframe_b200.multigpu import rank_frames
    spp_total = args.c5_spp
    if rank == 0:
        ensure_pack("c4_stress")
    barrier()
    pack = lf.ScenePack(ensure_pack("c4_stress"))
    pt = lf.PathTracer(local_rank)
    pt.set_stream(stream.cuda_stream)
    pt.upload_pack(pack)
    W, H = pt.params.width, pt.params.height
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pt.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        pt.nccl_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    f0, n, st = rank_frames(2, spp_total, rank, world)
    chunk = 32
    pt.render_frames(f0, min(chunk, n), st)            # warm-up (kernels, clocks), then the collective
    if world > 1:
        pt.reduce()
    torch.cuda.synchronize()
    pt.clear()
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    done = 0
    while done < n:
        m = min(chunk, n - done)
        pt.render_frames(f0 + done * st, m, st)
        done += m
    if world > 1:
        pt.reduce()
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    img = pt.read_accum() if rank == 0 else None
    pt.close()
    rec = {"workload": WORKLOADS["c4_stress"][1], "resolution": [W, H], "spp_total": spp_total, "spp_per_rank": n, "n_gpus": world, "scaling": "strong",
           "seconds": ms * 1e-3, "value": W * H * spp_total / (ms * 1e-3), "unit": "samples/s",
           "collective": "lfcuda_reduce (ncclAllReduce, fp32 sum of the 3840x2160x3 accumulation buffer) inside the timed region" if world > 1 else "none (single GPU)"}
    if rank == 0:
        m = img / np.float32(spp_total)
        rec["mean_rgb"] = [float(x) for x in m.mean(axis=(0, 1))]
        rec["finite"] = bool(np.isfinite(img).all())
    return rec


if __name__ == "__main__":
    main()
