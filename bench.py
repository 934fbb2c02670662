#!/usr/bin/env python
"""Benchmark of the IPDM domain-progressive inference path (BASELINE.json metric: slices/sec).

    python bench.py --gpus N --steps K --warmup W            # this build (libipdm_b200.so), one rank per GPU
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

A "step" is one pass of the whole hot path over one batch of synthetic slices per GPU:
projection-domain guided process (t_start_proj) -> FBP -> sharpen -> image-domain guided process
(t_start_img) -> "ultra" pass [5,5,5], i.e. `progressive_denoiser` of the reference with
convertor="FBP", ultra_img_denoise=True on 2000x912 sinograms -> 512x512 images.
`value` is timed with inputs resident in HBM; `e2e` goes through the reference-facing API
(`progressive_domain_denoiser.data_sample_load` + `.progressive_denoiser`) from pinned host buffers
to a host result.  Slices are sharded across ranks with no collective on the data path (weak
scaling); the only collective is the final gather of the [B,1,512,512] results.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "ipdm-pytorch_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

HBM_FALLBACK_GBS, TENSOR_FALLBACK_TFLOPS = 6650.0, 1590.0        # /opt/skills/guides/B200_PROFILING.md
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from an `ncu --set full` capture
CONV_TC_NCU_TRAFFIC = dict(bytes_per_launch=2.297e9,
                           note="profiles/r01_conv_halo_persistent_ncu_full.md: conv_halo_persistent_kernel<128,6>, bf16, 16 x 500x228, 128->128 3x3 "
                                "+residual: 1.401 GB read + 0.896 GB written per launch; algorithmic bytes of that layer 2.33 GB (bf16 operand 0.47 + fp32 "
                                "residual 0.93 + fp32 output 0.93)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="slices per GPU per step (BASELINE.json configs[2]: batch 16 on one B200)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="UNet operand precision (configs[2]: bf16 UNet, fp32 sampler state); tf32 / fp32 are the parity modes")
    ap.add_argument("--t_start_proj", type=int, nargs="+", default=[15, 15, 15])
    ap.add_argument("--t_start_img", type=int, nargs="+", default=[15, 15, 15])
    ap.add_argument("--skip_cpu_baseline", action="store_true")
    ap.add_argument("--cuda_graph", type=int, default=0, help="1: capture the whole progressive pass in one CUDA graph")
    return ap.parse_args()


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), which="measured (MEASURED_PEAKS.json; sustained bf16)")
    return dict(hbm=HBM_FALLBACK_GBS, tensor=TENSOR_FALLBACK_TFLOPS, which="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# -------------------------------------------------------------------------------------------------
# CPU legs (oracle port): the reference algorithm on the host cores, bounded sample
# -------------------------------------------------------------------------------------------------
def cpu_reference_sample(t_start_proj, t_start_img):
    """Times one proj-UNet forward, one img-UNet forward, one FBP and one sampler step of the oracle on the host and
    extrapolates one slice: n_proj*t_proj + n_img*t_img + t_fbp + steps*t_step (the reference is serial per slice)."""
    import numpy as np
    import torch
    from oracle import ipdm_oracle as O
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    torch.manual_seed(0)
    pnet, inet = O.UNetOracle(**O.PROJ_UNET).eval(), O.UNetOracle(**O.IMG_UNET).eval()
    g = torch.Generator().manual_seed(0)
    xp, xi = 3 * torch.rand(1, 1, 2000, 912, generator=g), 0.2 * torch.rand(1, 1, 512, 512, generator=g)
    t0 = time.perf_counter(); ep = pnet(xp, torch.full((1,), 7, dtype=torch.long)); t_proj = time.perf_counter() - t0
    t0 = time.perf_counter(); ei = inet(xi, torch.full((1,), 7, dtype=torch.long)); t_img = time.perf_counter() - t0
    tab = O.Tables(1000, 5)
    t0 = time.perf_counter(); O.p_sample_condition(tab, ep, xp, xp, 7, 0.4, False, torch.randn(xp.shape, generator=g)); t_sp = time.perf_counter() - t0
    t0 = time.perf_counter(); O.p_sample_condition(tab, ei, xi, xi, 7, 0.45, True, torch.randn(xi.shape, generator=g)); t_si = time.perf_counter() - t0
    t0 = time.perf_counter(); O.fbp_convert(xp[:, 0].numpy()); t_fbp = time.perf_counter() - t0
    n_proj, n_img = sum(t_start_proj), sum(t_start_img) + 15
    per_slice = n_proj * (t_proj + t_sp) + n_img * (t_img + t_si) + t_fbp
    return dict(per_slice_s=per_slice, t_proj=t_proj, t_img=t_img, t_fbp=t_fbp, cores=torch.get_num_threads(),
                sample=f"1 proj UNet fwd 2000x912 ({t_proj:.2f}s) + 1 img UNet fwd 512x512 ({t_img:.2f}s) + 1 sampler step each "
                       f"+ 1 FBP ({t_fbp:.2f}s, C oracle, OpenMP); extrapolated to {n_proj}+{n_img} forwards per slice")


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(args.t_start_proj, args.t_start_img)
        if i >= args.warmup:
            vals.append(last["per_slice_s"])
        if i == 0 and last["per_slice_s"] > 0 and (args.warmup + args.steps) * 25 > 600:
            pass
    per_slice = sum(vals) / len(vals)
    v = 1.0 / per_slice
    line = dict(metric="ipdm_progressive_slices_per_sec", value=v, unit="slices/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per_slice * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference",
                config=dict(workload=workload_name(args, 1), note="reference algorithm (oracle port: torch CPU UNet + C FBP) on the host; "
                            "the Python reference itself cannot travel to the GPU box; one slice at a time (the reference cannot batch)"),
                cpu_baseline=dict(value=v, unit="slices/s", cores=last["cores"], kind="port", sample=last["sample"]),
                e2e=dict(value=v, unit="slices/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def workload_name(args, batch):
    return (f"IPDM progressive inference, convertor=FBP, t_start_proj={args.t_start_proj}, t_start_img={args.t_start_img} + ultra [5,5,5], "
            f"{batch} slice(s)/GPU/step, sinogram 2000x912 -> image 512x512, random-init UNets (28.4M + 29.1M params)")


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from Config.default_config import default_cfg
    from ipdm_pytorch_b200 import engine, synthetic
    from Utils.train_test_utils import progressive_domain_denoiser

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    import contextlib
    import io
    import tempfile
    with contextlib.redirect_stdout(io.StringIO()):          # the reference's option loader prints; stdout carries ONE JSON line
        opt = default_cfg(["--load_option_path", os.path.join(PKG, "Config/Mayo-Config/test_progressive_option.json"), "--device", f"cuda:{local_rank}"])
    opt.load_img_model_path = opt.load_proj_model_path = None
    for k in ("test_dataset_path_FD_img", "test_dataset_path_LD_img", "test_dataset_path_FD_proj", "test_dataset_path_LD_proj"):
        setattr(opt, k, None)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = progressive_domain_denoiser(opt, result_save_path=tempfile.mkdtemp(prefix="ipdm_bench_"))
        model.update_opt(dict(convertor="FBP", save_it_state_img=False, save_it_state_proj=False, ultra_img_denoise=True,
                              t_start_proj=args.t_start_proj, t_start_img=args.t_start_img, precision=args.precision, noise_seed=1234 + rank,
                              cuda_graph=bool(args.cuda_graph)))
    # synthetic slices of this rank's shard (weak scaling: B per GPU), pinned on the host
    host = torch.from_numpy(synthetic.cheap_sinogram(B, seed=100 + rank))[:, None].contiguous().pin_memory()
    host_out = torch.empty(B, 1, 512, 512).pin_memory()
    resident = host.to(dev)

    def step_resident():
        model.ldproj = resident
        return model.progressive_denoiser()

    def step_e2e():
        model.ldproj = host.to(dev, non_blocking=True)
        out = model.progressive_denoiser()
        host_out.copy_(out, non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, out

    engine.launch_count_reset()
    for w_i in range(args.warmup):
        step_resident()
        if w_i == 0:
            launches_first_step = engine.launch_count()          # graph mode: the kernels recorded by the capture pass = kernels per replay
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    engine.launch_count_reset()
    ms, out = timed(step_resident, args.steps)
    launches = engine.launch_count()
    if args.cuda_graph and launches == 0 and args.warmup > 0:
        launches = launches_first_step * args.steps                  # replays do not pass through the host-side launch counter
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:                                   # the only collective: final gather of the results
        from ipdm_pytorch_b200.sharding import gather_slices
        gathered = gather_slices(out.contiguous(), world * B)
        assert gathered.shape[0] == world * B
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    if rank != 0:
        return
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    # per-family profile of ONE extra step (CUDA events around every launch, on the launching stream)
    engine.profile_enable(True)
    step_resident()
    prof = engine.profile_collect()
    engine.profile_enable(False)
    pk = peaks()
    tc_ms, tc_flops, tc_n = prof["conv_halo_persistent"]                  # the dominant kernel of the step
    fam_ms = tc_ms + prof["conv_tc"][0]
    fam_flops = tc_flops + prof["conv_tc"][1]
    total_prof_ms = sum(v[0] for v in prof.values())
    kind = {"bf16": "kind::f16 (bf16 operands, fp32 accumulate)", "tf32": "kind::tf32", "fp32": "kind::tf32 x3 (3xTF32 split)"}[args.precision]
    if args.precision == "fp32" or not tc_ms:                              # the 3xTF32 mode runs the one-tile kernel: report the family
        tc_ms, tc_flops, tc_n = fam_ms, fam_flops, tc_n + prof["conv_tc"][2]
        kname = f"conv_tc_kernel<N,S,SPLIT> (tcgen05 {kind} implicit-GEMM conv)"
    else:
        kname = f"conv_halo_persistent_kernel<N,NB> (tcgen05 {kind} implicit-GEMM 3x3 conv, halo reuse, persistent)"
    roof = dict(bound="tensor", kernel=kname, achieved=tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                peak=pk["tensor"], unit="TFLOP/s", traffic=CONV_TC_NCU_TRAFFIC["bytes_per_launch"] if args.precision == "bf16" else None,
                traffic_note=CONV_TC_NCU_TRAFFIC["note"], peak_source=pk["which"],
                note="achieved = layer FLOPs (2*pixels*taps*K*C_out with K padded to the operand stride; discarded tile columns not counted) of every launch of this kernel in one profiled step / their summed "
                     "CUDA-event time on the launching stream; peak is dense bf16 (a kind::tf32 MMA runs at half that rate, so 0.5 is the "
                     "ceiling in tf32 mode).  conv_tc_family_* = the same over ALL tensor-core conv kernels (adds 1x1 / stride-2 / N=16 / qkv "
                     "layers, most of which are HBM-bound)",
                share_of_step=tc_ms / total_prof_ms if total_prof_ms else None, launches_per_step=tc_n,
                conv_tc_family_achieved=fam_flops / (fam_ms * 1e-3) / 1e12 if fam_ms else None,
                conv_tc_family_share_of_step=fam_ms / total_prof_ms if total_prof_ms else None)
    roof["frac"] = roof["achieved"] / roof["peak"] if roof["achieved"] else None
    families = {}
    for k, (m, w, n) in prof.items():
        unit = "TFLOP/s" if k in ("conv_tc", "conv_halo_persistent", "attention") else "GB/s"
        rate = (w / (m * 1e-3) / (1e12 if unit == "TFLOP/s" else 1e9)) if m else None
        families[k] = dict(ms=round(m, 3), launches=n, rate=None if rate is None else round(rate, 2), unit=unit,
                           frac_of_peak=None if rate is None else round(rate / (pk["tensor"] if unit == "TFLOP/s" else pk["hbm"]), 4))
    p_flops = model.proj_model.cuda_handle().flops(B, 2000, 912)
    i_flops = model.img_model.cuda_handle().flops(B, 512, 512)
    step_flops = sum(args.t_start_proj) * p_flops + (sum(args.t_start_img) + 15) * i_flops
    line = dict(metric="ipdm_progressive_slices_per_sec", value=value, unit="slices/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"tf32": "tf32 (fp32 state, tcgen05 kind::tf32, fp32 accumulate)", "fp32": "f32 (3xTF32)", "bf16": "bf16"}[args.precision],
                data="synthetic",
                config=dict(workload=workload_name(args, B), global_batch=world * B, parallelism=f"slice-sharded x{world}, no data-path collective",
                            l2="working set (activation arena of several GB per step) >> 126 MB L2; no explicit flush needed",
                            noise="in-kernel Philox4x32-10", cuda_graph=bool(args.cuda_graph), unet_tflop_per_step=step_flops / 1e12),
                e2e=dict(value=e2e, unit="slices/s", h2d_bytes_per_step=int(host.numel() * 4), d2h_bytes_per_step=int(host_out.numel() * 4),
                         ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), clocks=clocks, roofline=roof, kernel_families=families,
                unet_effective_tflops=step_flops * args.steps / (ms * 1e-3) / 1e12)
    line["fbp_batch64"] = fbp_batch64(dev, pk)
    if not args.skip_cpu_baseline and world == 1:             # rank 0 at N=1 only
        c = cpu_reference_sample(args.t_start_proj, args.t_start_img)
        line["cpu_baseline"] = dict(value=1.0 / c["per_slice_s"], unit="slices/s", cores=c["cores"], kind="port", sample=c["sample"])
    print(json.dumps(line))


def fbp_batch64(dev, pk):
    """BASELINE.json configs[1]: the FBP convertor alone on 64 sinograms (fan-beam weight + ramp filter + pixel-driven backprojection).
    Algorithmic HBM bytes per slice: sinogram read + filtered write + filtered read + image write; the backprojection itself is bound
    by the 2000 x 512 x 512 pixel-view updates per slice (gather + interpolate from L1/shared), not by HBM."""
    import torch
    from ipdm_pytorch_b200 import engine, synthetic
    n = 64
    sino = torch.from_numpy(synthetic.cheap_sinogram(4, seed=7)).to(dev).repeat(n // 4, 1, 1).contiguous()
    plan = engine.FBPPlan(max_batch=n)
    out = torch.empty(n, 512, 512, device=dev)
    for _ in range(2):
        plan.forward(sino, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        plan.forward(sino, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    bytes_alg = n * (3 * 2000 * 912 * 4 + 512 * 512 * 4)
    return dict(workload="FBP convertor alone, 64 sinograms 2000x912 -> 512x512", ms=round(ms, 3), slices_per_s=round(n / (ms * 1e-3), 1),
                algorithmic_gb_s=round(bytes_alg / (ms * 1e-3) / 1e9, 1), frac_of_hbm=round(bytes_alg / (ms * 1e-3) / 1e9 / pk["hbm"], 4),
                pixel_view_updates_per_s=round(n * 2000 * 512 * 512 / (ms * 1e-3) / 1e9, 1), updates_unit="G/s")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if rank == 0:
            # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to run on
            os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
            run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
