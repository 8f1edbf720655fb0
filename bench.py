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
def oracle_sample_rate(pack_path, budget_s, tile=(320, 180)):
    """Oracle port on a bounded sample: the centre tile of the frame, as many spp as fit the time budget."""
    from oracle_api import Oracle
    o = Oracle(pack_path)
    W, H = o.params.width, o.params.height
    tw, th = min(tile[0], W), min(tile[1], H)
    while W % tw:
        tw -= 1
    while H % th:
        th -= 1
    o.update_params(tile_width=tw, tile_height=th)
    tx, ty = (W // tw) // 2, (H // th) // 2
    t0 = time.time()
    o.render_frames(2, 1, 1, tx, ty)
    t1 = time.time() - t0
    spp = int(max(2, min(512, budget_s / max(t1, 1e-3))))
    t0 = time.time()
    o.render_frames(3, spp, 1, tx, ty)
    dt = time.time() - t0
    o.close()
    return tw * th * spp / dt, f"centre tile {tw}x{th} of the {W}x{H} frame, {spp} spp, frames 3..{spp + 2} ({tw * th * spp} pixel-samples, {dt:.1f} s)"


def run_reference(args, workload, desc):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    pack = ensure_pack(WORKLOADS[workload][0])
    cores = os.cpu_count()
    from oracle_api import Oracle
    o = Oracle(pack)
    W, H = o.params.width, o.params.height
    tw, th = 320, 180
    while W % tw:
        tw -= 1
    while H % th:
        th -= 1
    o.update_params(tile_width=tw, tile_height=th)
    tx, ty = (W // tw) // 2, (H // th) // 2
    t0 = time.time(); o.render_frames(2, 1, 1, tx, ty); t1 = time.time() - t0
    # size one step so that warmup + steps finish within ~2.5 minutes
    per_step_budget = min(2.0, 150.0 / max(1, args.steps + args.warmup))
    spp = int(max(1, min(256, per_step_budget / max(t1, 1e-3))))
    frame = 3
    for _ in range(args.warmup):
        o.render_frames(frame, spp, 1, tx, ty); frame += spp
    t0 = time.time()
    for _ in range(args.steps):
        o.render_frames(frame, spp, 1, tx, ty); frame += spp
    dt = time.time() - t0
    o.close()
    value = tw * th * spp * args.steps / dt
    sample = f"oracle port (CPU restatement of the GLSL, OpenMP): centre tile {tw}x{th} of the {W}x{H} frame, {spp} spp per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "spp_per_step": spp, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
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
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "reference", "sample": what}
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
        return {"kind": "reference", "workload": what, "samples_per_s_steady": j["samples_per_s_steady"], "first_step_s": j["first_step_s"],
                "render_s": j["render_s"], "cores": os.cpu_count(), "lp_num_threads": j["lp_num_threads"], "gl": j["gl_renderer"] + " / " + j["gl_version"]}
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
    if world > 1:
        tms = torch.tensor([ms], device="cuda"); dist.all_reduce(tms, op=dist.ReduceOp.MAX); ms = float(tms.item())
        tl = torch.tensor([launches], device="cuda", dtype=torch.int64); dist.all_reduce(tl); launches = int(tl.item())
    total_samples = npix * S * K * world
    value = total_samples / (ms * 1e-3)

    # ---- e2e through the C ABI with host buffers: uniforms in, accumulated image out, every step
    host_img = torch.empty((H, W, 3), dtype=torch.float32).pin_memory().numpy()
    h2d = ctypes.sizeof(lf.LfParams) + ctypes.sizeof(lf.LfCamera)
    d2h = npix * 3 * 4 if rank == 0 else 0
    pt.clear()
    barrier(); sync()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(K):
        f0, st = step_frames(Wm + K + i)
        pt.set_params(pt.params); pt.set_camera(pack.camera())       # host structs -> device uniforms
        pt.render_frames(f0, S, st)
        if world > 1:
            # out-of-place sum for read-out: reduce a snapshot so that the local accumulator keeps only local samples
            ptr, n = pt.accum_device_ptr()
            snap = torch.as_tensor(_DevBuf(ptr, n), device="cuda").clone()
            dist.all_reduce(snap)
            if rank == 0:
                torch.from_numpy(host_img.reshape(-1)).copy_(snap, non_blocking=False)
        else:
            pt.read_accum(host_img)                                   # D2H + synchronize
    e1.record(stream)
    sync(); barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        tms = torch.tensor([e2e_ms], device="cuda"); dist.all_reduce(tms, op=dist.ReduceOp.MAX); e2e_ms = float(tms.item())
    e2e_value = total_samples / (e2e_ms * 1e-3)

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
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic = None
    prof = os.path.join(ROOT, "profiles", "extend_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    rays = c["rays_closest"] + c["rays_shadow"]
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "launches": dom_launches, "ms_per_launch": dom_ms / dom_launches, "bytes_per_launch": dom_bytes / dom_launches,
        "bytes_per_ray": (ext_bytes + sh_bytes) / max(1, rays),
        "stage_ms": {k: round(v["ms"], 3) for k, v in stage.items()},
        "shadow_achieved_gbs": sh_bytes * scale / (stage["shadow"]["ms"] * 1e-3) / 1e9 if stage["shadow"]["ms"] > 0 else None,
    }
    if rank == 0:   # the L2 denominator north_star asks for (not in MEASURED_PEAKS.json): measured here, 16 MiB working set
        try:
            l2 = pt.measure_read_bandwidth(16 << 20, 10)
            hbm_read = pt.measure_read_bandwidth(2 << 30, 3)
            roofline["l2_read_gbs_measured"] = l2
            roofline["hbm_read_gbs_measured"] = hbm_read
            roofline["frac_of_l2"] = achieved / l2
        except Exception as e:
            roofline["l2_read_gbs_measured"] = f"failed: {e}"
    # the bound that ncu shows to be the active one (l1tex data-pipe wavefronts): the rate at which this GPU delivers
    # divergent 64-byte node fetches that hit L2, measured by tools/l1_probe.cu and recorded under profiles/
    ceil_path = os.path.join(ROOT, "profiles", "l1_fetch_ceiling.json")
    if os.path.exists(ceil_path):
        try:
            ceil = json.load(open(ceil_path))["node_fetch_l2_hit_gbs"]
            roofline["l1_node_fetch_ceiling_gbs"] = ceil
            roofline["frac_of_l1_node_fetch_ceiling"] = achieved / ceil
        except Exception:
            pass
    mrays = rays * scale * world / (ms * 1e-3) / 1e6

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "resolution": [W, H], "spp_per_step": S, "spp_total": S * K * world, "full_config_spp": full_spp,
                   "max_depth": pt.params.max_depth, "parallelism": f"spp-split x{world}" if world > 1 else "single GPU",
                   "kernels": "megakernel" if args.kernel_mode == 1 else "wavefront",
                   "l2_policy": "scene geometry (~165 MB) + 4 M-path state (~1 GB) exceed the 126 MB L2 every step; no flush needed"},
        "mrays_per_s": mrays, "rays_per_sample": rays / max(1, c["samples"]),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K,
                "path": "lfcuda_set_params + lfcuda_set_camera + lfcuda_render_frames + lfcuda_read_accum (host buffers)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, sample = oracle_sample_rate(pack_path, 12.0)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": sample}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    pt.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
