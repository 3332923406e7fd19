#!/usr/bin/env python
"""bench.py -- decode tokens/sec of a LLaMA-7B-shaped model over the quantised KV cache (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one batch-1 decode step at cache length L of the named workload: for every layer q/k/v projection,
fused device-side append (NUQ quantise + top-K outlier split + pack), fused attend over the packed cache
(LUT dequant + outlier SpMV + RoPE + softmax + V), o_proj and MLP, then norm + lm_head.  Weights are random-init
fp16 of the LLaMA architecture, caches are filled with synthetic K/V through the real prefill packers.
The whole step is one CUDA graph; replays re-run the step at the SAME cache length (the fused append overwrites
its slot), so all K timed steps are measured at the named seqlen.

N > 1, one process per GPU under torch.distributed / NCCL, two layouts (--parallelism):
  pp  layer-group pipeline -- the reference's own multi-GPU scheme (modeling_llama.py:2428-2453) and what north_star
      prescribes: NCCL send/recv of the [hidden] fp16 vector over NVLink.  At batch 1 the stages run one after
      another, so it buys capacity (1M-token contexts), not tokens/sec.
  sp  sequence-sharded attention: weights replicated, every rank holds 1/N of every layer's
      cache, per layer one all-gather of the partial (out, lse) results (H*129 floats per rank) and a merge kernel.
      This is the layout in which decode speeds up with N; it is the default ("auto") when the workload divides.
Both are strong scaling of the same fixed workload ("scaling": "strong").

--impl reference: the reference's CPU implementation of the path, i.e. the C port of the kernel semantics
(oracle/kvq_oracle_port.c; /root/reference does not exist on the GPU box and the reference's CPU path is Python)
timed on the host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, bits, L (quantised slots), n_sink, description = BASELINE.json configs[i])
    "7b-3b-128k": ("7b", 3, 131072, 5, "LLaMA-7B 3b NUQ + 1% outliers + 5 fp16 sink tokens, seqlen 128K, 1xB200 (configs[2])"),
    "7b-4b-128k": ("7b", 4, 131072, 0, "LLaMA-7B 4b NUQ + 1% outliers, seqlen 128K (north_star roofline target)"),
    "7b-4b-32k": ("7b", 4, 32768, 0, "LLaMA-7B 4b NUQ + 1% outliers, seqlen 32K, 1xB200 decode (configs[1])"),
    "7b-4b-4k": ("7b", 4, 4096, 0, "single-node smoke size"),
}
DEFAULT_WORKLOAD = "7b-3b-128k"   # BASELINE.json's metric is quoted at seqlen 128K; this config fits one GPU


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def build_quantizer(cfg_bits, H, device):
    from kvquant_b200 import synth, cache as kc
    sp = synth.SynthSpec(H, 128, seed=0)
    cal = synth.calibrate(sp, cfg_bits, calib_tokens=512, seed=7)
    t = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], H, device=device)
    klut = dict(lut=t["lut"], lut2=None, thr_lower=t["thr_lower"], thr_upper=t["thr_upper"])
    return sp, dict(klut=klut, v_cent=cal["v"][2][0])


# ------------------------------------------------------------------------------------------------------------
# reference arm: CPU port of the path on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_baseline_run(layer_arrays, bits, H, Lmax, L, n_out, n_layers, theta, pos_offset, budget_s=12.0):
    """Time the C port's attend for one layer over a token sample sized to ~budget_s of CPU work; extrapolate to
    tokens/sec of the whole model's attention (n_layers x L tokens; the dense GEMVs are NOT added, which only
    favours the CPU number).  Returns (tokens_per_sec, cores, sample description)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_oracle_c
    lib = build_oracle_c.load()
    lib.kvq_port_set_threads(len(os.sched_getaffinity(0)))   # not OMP_NUM_THREADS (torchrun sets it to 1)
    cores = lib.kvq_port_threads()
    a = layer_arrays
    q = np.ascontiguousarray(a["q"], dtype=np.float32)
    out = np.zeros((H, 128), np.float32)

    def run(Ls):
        scratch = np.zeros((H, Ls), np.float32)
        t0 = time.perf_counter()
        lib.kvq_port_attend(bits, q.ctypes.data, a["kcache"].ctypes.data, a["klut"].ctypes.data,
                            a["k_out"].ctypes.data, a["k_idx"].ctypes.data, a["vcache"].ctypes.data,
                            a["vlut"].ctypes.data, a["v_out"].ctypes.data, a["v_idx"].ctypes.data, n_out, H, Lmax, Ls,
                            theta, pos_offset, out.ctypes.data, scratch.ctypes.data)
        return time.perf_counter() - t0

    t_probe = run(min(L, 2048))
    per_tok = t_probe / min(L, 2048)
    Ls = int(max(2048, min(L, budget_s / 3 / per_tok)))
    ts = sorted(run(Ls) for _ in range(3))
    t_layer_full = ts[1] * (L / Ls)
    tok_s = 1.0 / (t_layer_full * n_layers)
    return tok_s, cores, "C port (oracle/kvq_oracle_port.c, OpenMP) of one layer's attend over %d of %d tokens, median of 3, " \
        "extrapolated to %d layers x %d tokens; GEMVs excluded" % (Ls, L, n_layers, L)


def ncu_traffic(bits, L):
    """dram__bytes_read.sum + dram__bytes_write.sum of one fused attend (all its kernels), per launch, from the
    committed ncu --set full capture of the same shape (profiles/ncu_traffic.json); None when no capture matches."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            for e in json.load(f)["entries"]:
                if e["bits"] == bits and abs(e["L"] - L) <= 1024:
                    return e["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="auto", choices=["auto", "pp", "sp"],
                    help="N>1: pp = layer-group pipeline (the reference's scheme, north_star: buys capacity, not "
                         "tokens/sec at batch 1); sp = sequence-sharded attention with replicated weights (SURVEY 8e-2 / "
                         "8f-1: the layout in which decode speeds up with N).  auto = sp when the workload fits, else pp")
    ap.add_argument("--graph", default="dynamic", choices=["dynamic", "static"],
                    help="dynamic: cache length / position live on the device and every replay is the NEXT decode step "
                         "(growing cache); static: every replay re-runs the step at the captured length")
    ap.add_argument("--sp-exchange", default="nccl", choices=["nccl", "p2p"],
                    help="sp only: how the per-GPU partial attention results meet -- nccl: all_gather + merge kernel "
                         "(default, validated); p2p: EXPERIMENTAL peer-memory exchange fused with the merge")
    ap.add_argument("--torch-profile", default="", help="write a per-kernel table of 3 graph replays to this file (diagnostic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import numpy as np
    import torch
    import torch.distributed as dist
    from kvquant_b200 import decode as kd, synth, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    model, bits, L, n_sink, desc = WORKLOADS[args.workload]
    metric = "decode tokens/sec @ seqlen %dK (LLaMA-7B, bs1)" % (L // 1024)
    base = {"metric": metric, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (random-init fp16 LLaMA-7B weights, synthetic K/V packed by the real prefill packers)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    par = args.parallelism
    if par == "auto":
        par = "sp" if (world > 1 and L % world == 0) else "pp"
    par_label = ("sp%d" if (par == "sp" and world > 1) else "pp%d") % world
    if args.impl == "reference":
        par = "pp"          # the CPU arm runs the whole job (all L tokens, all layers) on rank 0's host cores
    sp_mode = (par == "sp" and world > 1)
    if sp_mode and L % world:
        raise SystemExit("sp needs seq_len divisible by the number of GPUs")
    L_local = L // world if sp_mode else L
    # room for every step this run appends (warm-up + profile + timed, device and e2e loops)
    headroom = (2 * (args.steps + args.warmup) + 3 + 8 + 63) // 64 * 64
    cfg = kd.DecodeConfig.llama7b(bits=bits, n_sink=n_sink, max_len=L_local + headroom)
    sp, quantizer = build_quantizer(bits, cfg.n_heads, dev)
    config = {"workload": args.workload, "description": desc, "bits": bits, "seq_len": L + n_sink, "n_sink": n_sink,
              "outliers": "1% (21+21 per token per cache)", "layers": cfg.n_layers, "parallelism": par_label,
              "l2_policy": "inputs larger than L2: every step streams all layers' caches (>= 4 GB) and 13.5 GB of weights",
              "step": ("one CUDA-graph replay = the next decode step of a growing cache (length and position live in "
                       "device memory; step i appends slot L+i and attends over L+i+1 slots)") if args.graph == "dynamic"
              else "one CUDA-graph replay of a full decode step at fixed cache length"}

    # ---------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # build ONE layer's cache on the GPU (real packers), bring it to the host, time the CPU port
        from kvquant_b200.cache import LayerCache
        lc = LayerCache.from_luts(bits, cfg.n_heads, cfg.max_len, quantizer["klut"], quantizer["v_cent"], device=dev)
        synth.fill_layer_cache_gpu(lc, sp, L, seed=0)
        arrs = dict(kcache=lc.kcache.cpu().numpy(), vcache=lc.vcache.cpu().numpy(), klut=lc.klut.cpu().numpy(),
                    vlut=lc.vlut.cpu().numpy(), k_out=lc.k_outliers.cpu().numpy(), k_idx=lc.k_outlier_idx.cpu().numpy(),
                    v_out=lc.v_outliers.cpu().numpy(), v_idx=lc.v_outlier_idx.cpu().numpy(), q=sp.q_vec(1))
        vals = []
        info = None
        # each step is a bounded sample; the whole --steps/--warmup run stays within ~2.5 minutes of CPU time
        ref_budget = max(1.5, min(6.0, 150.0 / max(1, args.warmup + args.steps)))
        for i in range(args.warmup + args.steps):
            tok_s, cores, sample = cpu_baseline_run(arrs, bits, cfg.n_heads, cfg.max_len, L, lc.n_out, cfg.n_layers,
                                                    cfg.rope_theta, n_sink, budget_s=ref_budget)
            if i >= args.warmup:
                vals.append(tok_s)
            info = (cores, sample)
        v = float(np.median(vals))
        line = dict(base, impl="reference", value=v, ms_per_step=1000.0 / v, config=config, gpu_launches=0,
                    cpu_baseline={"value": v, "unit": "tokens/s", "cores": info[0], "kind": "port", "sample": info[1]},
                    e2e={"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    clocks=None, roofline=None)
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------------------------------------------------
    if sp_mode:
        lo, hi = 0, cfg.n_layers
        stage = kd.DecoderStage(cfg, lo, hi, dev, quantizer, seed=0, with_head=True, sp=(rank, world))
        stage.global_pos = n_sink + L
        if args.sp_exchange == "p2p":
            from kvquant_b200.p2p import PeerExchange
            stage.xchg = PeerExchange(rank, world, cfg.n_heads, dev)
            config["sp_exchange"] = "p2p (experimental)"
    else:
        lo, hi = kd.partition_layers(cfg.n_layers, world, rank)
        stage = kd.DecoderStage(cfg, lo, hi, dev, quantizer, seed=0, with_head=(rank == 0))
    t_fill = time.time()
    for i, ly in enumerate(stage.layers):
        if sp_mode:
            ly.cache.pos_base = rank * L_local
        synth.fill_layer_cache_gpu(ly.cache, sp, L_local, seed=(lo + i) * 16 + (rank if sp_mode else 0))
    torch.cuda.synchronize()
    t_fill = time.time() - t_fill

    n0 = _lib.launch_count()
    gs = kd.GraphedStage(stage, L_local, first=(rank == 0 or sp_mode), last_to_logits=(world == 1 or sp_mode),
                         dynamic=(args.graph == "dynamic"), pos=(n_sink + L) if sp_mode else None)
    launches_per_step = (_lib.launch_count() - n0) // (5 if sp_mode else 3)   # eager warm-up passes + 1 capture pass
    if world > 1 and rank == 0:
        head_graph_in = torch.zeros(cfg.hidden, dtype=torch.float16, device=dev)
    pinned_tok = torch.zeros(1, dtype=torch.long).pin_memory()
    pinned_logits = torch.zeros(cfg.vocab, dtype=torch.float16).pin_memory()
    logits_dev = [None]

    def step_device():
        """one decode step, inputs resident on the device"""
        if world == 1 or sp_mode:
            gs.replay()
            logits_dev[0] = gs.logits
            return
        if rank > 0:
            dist.recv(gs.x_in, src=rank - 1)
        gs.replay()
        dist.send(gs.y, dst=(rank + 1) % world)
        if rank == 0:
            dist.recv(head_graph_in, src=world - 1)
            logits_dev[0] = stage.head(head_graph_in)

    def step_e2e(i):
        """same step through host buffers: token id H2D from pinned memory, logits D2H to pinned memory"""
        if rank == 0 or sp_mode:
            pinned_tok[0] = (17 * i + 3) % cfg.vocab
            gs.tok.copy_(pinned_tok, non_blocking=True)
        step_device()
        if rank == 0 or sp_mode:
            pinned_logits.copy_(logits_dev[0], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step_device()
    if args.torch_profile and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(3):
                step_device()
            torch.cuda.synchronize()
        with open(args.torch_profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda i: step_device(), args.steps)
    for i in range(args.warmup):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = 1000.0 / ms_step
    e2e_value = 1000.0 / (ms_e2e / args.steps)

    # ---- roofline of the dominant op (fused attend), measured live on this stream, cycling over the layers -------
    roof = None
    cpu_b = None
    if rank == 0:
        peak, peak_src = measured_peak()
        layers = stage.layers
        q = torch.randn((cfg.n_heads, 128), device=dev).half().float()
        reps = max(2, 64 // len(layers))

        def time_loop(fn):
            for ly in layers[:2]:
                fn(ly)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                for ly in layers:
                    fn(ly)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / (reps * len(layers))

        Lq = layers[0].cache.len      # current length (the timed steps appended to the cache)
        ms_att = time_loop(lambda ly: ly.cache.attend(q, rope_theta=cfg.rope_theta))
        from kvquant_b200 import quant_cuda as qc
        mulK = torch.zeros((1, cfg.n_heads, Lq), device=dev)
        pV = torch.softmax(torch.randn((1, cfg.n_heads, Lq), device=dev), -1)
        mulV = torch.zeros((1, cfg.n_heads, 128), device=dev)
        q1 = q[None].contiguous()
        kop = getattr(qc, "vecquant%dmatmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2" % bits)
        vop = getattr(qc, "vecquant%dmatmul_nuq_perchannel_transposed_mha_batched_fused_opt2" % bits)
        ms_k = time_loop(lambda ly: kop(q1, ly.cache.kcache, mulK, ly.cache.klut.view(cfg.n_heads, 128, -1), Lq,
                                        ly.cache.k_outliers, ly.cache.k_outlier_idx, cfg.rope_theta, n_sink))
        ms_v = time_loop(lambda ly: vop(pV, ly.cache.vcache, mulV, ly.cache.vlut, Lq, ly.cache.v_outliers,
                                        ly.cache.v_outlier_idx))
        n_out = layers[0].cache.n_out
        b_att = kd.layer_step_bytes(cfg, Lq)
        b_k = Lq * (cfg.hidden * bits // 8 + 8 * n_out)
        b_v = Lq * (cfg.hidden * bits // 8 + 8 * n_out + 4 * 2 ** bits)
        ach = b_att / ms_att / 1e6
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(bits, Lq),
                "peak_source": peak_src,
                "kernel": "kvq_attend = attend_init + k_outlier_pers + k_scores(3) + v_native + attend_combine",
                "algorithmic_bytes_per_launch": b_att, "ms_per_launch": ms_att,
                "per_kernel": {"k_scores_kernel": {"ms": ms_k, "bytes": b_k, "gbs": b_k / ms_k / 1e6, "frac": b_k / ms_k / 1e6 / peak},
                               "v_accum_kernel": {"ms": ms_v, "bytes": b_v, "gbs": b_v / ms_v / 1e6, "frac": b_v / ms_v / 1e6 / peak}},
                "attend_share_of_step": ms_att * len(layers) / ms_step}
        if not args.no_cpu_baseline:
            lc = layers[0].cache
            arrs = dict(kcache=lc.kcache.cpu().numpy(), vcache=lc.vcache.cpu().numpy(), klut=lc.klut.cpu().numpy(),
                        vlut=lc.vlut.cpu().numpy(), k_out=lc.k_outliers.cpu().numpy(), k_idx=lc.k_outlier_idx.cpu().numpy(),
                        v_out=lc.v_outliers.cpu().numpy(), v_idx=lc.v_outlier_idx.cpu().numpy(), q=q.cpu().numpy())
            tok_s, cores, sample = cpu_baseline_run(arrs, bits, cfg.n_heads, cfg.max_len, L_local, n_out,
                                                    cfg.n_layers * (world if sp_mode else 1), cfg.rope_theta, n_sink,
                                                    budget_s=15.0)
            cpu_b = {"value": tok_s, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        config.update(cache_fill_s=round(t_fill, 1), weight_bytes=stage.weight_bytes(),
                      cache_bytes_per_layer=kd.layer_step_bytes(cfg, L))
        line = dict(base, value=value, ms_per_step=ms_step, config=config, clocks=clocks,
                    e2e={"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 8,
                         "d2h_bytes_per_step": cfg.vocab * 2},
                    gpu_launches=int(launches_per_step * args.steps), roofline=roof, cpu_baseline=cpu_b)
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down: graphs that captured NCCL kernels must die before the communicator; a hung destroy must not hold
        # the job (the result line is already out), so leave through os._exit after a final barrier
        sys.stdout.flush()
        del gs
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
