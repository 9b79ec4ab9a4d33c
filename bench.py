#!/usr/bin/env python
"""Benchmark of the LLaMA-VQA training step (BASELINE.json metric: 7B NExT-QA train samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 7b-nextqa] [--impl ours|reference]

One JSON line on stdout (rank 0). A "step" = forward + hand-written backward + gradient all-reduce
(N>1) + AdamW update on the trainables, on a synthetic NExT-QA-shaped batch (B=8, S=128, F=10,
--vaq --qav) with random-init LLaMA-7B-shaped weights.
  value : whole-job samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API `model(data)` with HOST batch tensors: host planning,
          pinned H2D copy and a D2H read of the loss inside the timed region
  roofline     : tcgen05 GEMM launches sampled with CUDA events inside the timed region
  cpu_baseline : the oracle (CPU port of the reference's step): one real full-depth step at a bounded batch
  gpu_eager_baseline : the reference's op sequence in eager PyTorch on the same GPU (N = 1), outside the timed regions
  alt_build    : the same resident-batch measurement with the other operand build (bf16 when the default fp16 library is loaded), child
                 process on the same GPU (N = 1)
  dp_check / comm    : N > 1: hardware check that the reduced gradient is the rank mean; exposed NCCL wait and rank skew
`--impl reference` times that CPU port with all host threads, every timed step a real full-depth step at
a bounded batch (the reference itself is pure PyTorch and /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "7b-nextqa": dict(dim=4096, n_layers=32, n_heads=32, vocab_size=32000, multiple_of=256, adapter_layer=32, bsz=8, seqlen=128),
    "7b-dramaqa": dict(dim=4096, n_layers=32, n_heads=32, vocab_size=32000, multiple_of=256, adapter_layer=32, bsz=2, seqlen=384),
    "7b-tvqa": dict(dim=4096, n_layers=32, n_heads=32, vocab_size=32000, multiple_of=256, adapter_layer=32, bsz=1, seqlen=650),
    "13b-nextqa": dict(dim=5120, n_layers=40, n_heads=40, vocab_size=32000, multiple_of=256, adapter_layer=40, bsz=8, seqlen=128),
    "tiny": dict(dim=256, n_layers=4, n_heads=4, vocab_size=512, multiple_of=256, adapter_layer=4, bsz=8, seqlen=128),
}
WORKLOAD_NAMES = {
    "7b-nextqa": "LLaMA-7B NExT-QA-shaped max_seq_len=128 bs=8 max_feats=10 --vaq --qav",
    "7b-dramaqa": "LLaMA-7B DramaQA-shaped max_seq_len=384 bs=2 --vaq --qav",
    "7b-tvqa": "LLaMA-7B TVQA-shaped max_seq_len=650 bs=1 --sub --vaq --qav",
    "13b-nextqa": "LLaMA-13B NExT-QA-shaped max_seq_len=128 bs=8 --vaq --qav adapter_layer=40",
    "tiny": "tiny random-init LLaMA-VQA dim=256 4 layers",
}
MAX_FEATS, ADAPTER_LEN, BIAS, TAU = 10, 10, 3.5, 100.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")   # B200_PROFILING.md


def gemm_traffic():
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def flops_per_step(cfg, n_labelled, n_live=None, rows=None):
    """F_step = 3*(2*body + 3.5*attn) + 2*sum(head) + small; heads counted on the
    labelled rows actually evaluated (never count work not done). `n_live`: rows the LAST layer's wo / FFN
    actually process (only the rows the losses read); the rows it skips are subtracted, forward and backward.
    `rows`: rows of all three streams the row-wise GEMMs actually process (padding-free leg), default 3*B*S."""
    d, L, V, B, S = cfg["dim"], cfg["adapter_layer"], cfg["vocab_size"], cfg["bsz"], cfg["seqlen"]
    from flipped_vqa_b200.synthetic import ffn_hidden_dim
    hid = ffn_hidden_dim(d, cfg["multiple_of"])
    rows = 3 * B * S if rows is None else rows
    body3 = L * 2 * rows * (4 * d * d + 3 * d * hid)           # all three streams
    attn = L * 4 * B * d * (S * (S + 1) / 2 + S * ADAPTER_LEN)
    head = 2 * n_labelled * d * V
    small = L * 2 * ADAPTER_LEN * 2 * d * d + 2 * B * MAX_FEATS * 768 * d
    total = 2 * body3 + 3 * 3.5 * attn + 2 * head + small
    if n_live is not None:
        total -= 2 * (rows - n_live) * 2 * (d * d + 3 * d * hid)
    return total


def make_args():
    return argparse.Namespace(max_feats=MAX_FEATS, bias=BIAS, tau=TAU, llama_model_path="x/", audio=False, audio_only=False,
                              audio_merge="none", debug=False, vaq=True, qav=True, is_generation_task=False)


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.power, self.reasons = [], [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        # NVML is initialised ONCE per process, before the warm-up steps: doing it here, between the warm-up and the timed
        # region, left the GPU idle for ~0.2 s, after which it boosted past the power cap on the first timed step and was
        # throttled hard on the second (one 94 ms step among 74 ms ones; visible in step_ms_rank0)
        self.nv, self.h, self.max_mhz = ClockSampler._nvml(index)

    _cache = {}

    @staticmethod
    def _nvml(index):
        if index not in ClockSampler._cache:
            try:
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
                # every query the sampling thread will issue, once, before the warm-up: NVML initialises each lazily, and a first call
                # landing inside the timed region stalls that rank's launches (one 150 ms step among 75 ms ones -> every rank waits)
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetPowerUsage(h)
                (pynvml.nvmlDeviceGetCurrentClocksEventReasons if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons")
                 else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(h)
                ClockSampler._cache[index] = (pynvml, h, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            except Exception:
                ClockSampler._cache[index] = (None, None, None)
        return ClockSampler._cache[index]

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        pw = sorted(self.power)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    power_w=(round(pw[len(pw) // 2], 1) if pw else None))


# ---------------------------------------------------------------------------------------------------
class CpuOracleStep:
    """The ORACLE (CPU port of the reference's step, `llama/model.py:250-365`, fp32, autograd) at FULL DEPTH on a bounded batch.
    Timing only: every layer aliases ONE random layer's frozen weights (same shapes, FLOPs and bytes per layer; 0.8 GB instead of
    27 GB of host memory and no minute-long random fill), trainables are per layer as in the model."""

    def __init__(self, cfg, n_layers, threads):
        from oracle import llama_vqa_oracle as O
        from flipped_vqa_b200.synthetic import ffn_hidden_dim
        torch.set_num_threads(threads)
        self.O, self.cfg, self.L = O, cfg, n_layers
        d, V, H, S = cfg["dim"], cfg["vocab_size"], cfg["n_heads"], cfg["seqlen"]
        hid = ffn_hidden_dim(d, cfg["multiple_of"])
        g = torch.Generator().manual_seed(0)
        r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
        sd = {"tok_embeddings.weight": r(V, d), "output.weight": r(V, d), "norm.weight": torch.ones(d),
              "adapter_query.weight": r(ADAPTER_LEN * n_layers, d, std=1.0), "visual_proj.weight": r(d, 768, std=0.03),
              "temporal_emb.weight": r(MAX_FEATS, d, std=1.0)}
        shared = {"attention.wq.weight": r(d, d), "attention.wk.weight": r(d, d), "attention.wv.weight": r(d, d), "attention.wo.weight": r(d, d),
                  "feed_forward.w1.weight": r(hid, d), "feed_forward.w2.weight": r(d, hid), "feed_forward.w3.weight": r(hid, d),
                  "attention_norm.weight": torch.ones(d), "ffn_norm.weight": torch.ones(d)}
        for i in range(n_layers):
            p = f"layers.{i}."
            for k, v in shared.items():
                sd[p + k] = v
            sd[p + "attention.gate1"] = r(1, H, 1, 1, std=0.5)
            sd[p + "attention.gate2"] = torch.full((1, H, 1, 1), -BIAS)
        self.params = SimpleNamespace(dim=d, n_layers=n_layers, n_heads=H, vocab_size=V, norm_eps=1e-6, max_seq_len=S,
                                      adapter_len=ADAPTER_LEN, adapter_layer=n_layers)
        self.st = O.prepare_state(sd)
        self.names = O.trainable_names(self.st)

    def step(self, bsz, seed=1):
        from flipped_vqa_b200.synthetic import synthetic_batch
        data = synthetic_batch(bsz, self.cfg["seqlen"], self.cfg["vocab_size"], max_feats=MAX_FEATS, seed=seed)
        for n in self.names:
            self.st[n].grad = None
        t0 = time.perf_counter()
        losses = self.O.forward_losses(self.st, self.params, data, max_feats=MAX_FEATS, tau=TAU)
        sum(losses).backward()
        return time.perf_counter() - t0


def _bounded_batch(cfg, threads, n_steps, budget_s):
    """Largest batch in {B, B/2, ..., 1} whose `n_steps` full-depth CPU steps fit `budget_s`, estimated from a 2-layer probe at
    batch 1 (a discarded cold call first). Returns (batch, probe seconds per layer per sample)."""
    probe = CpuOracleStep(cfg, 2, threads)
    probe.step(1)
    t2 = probe.step(1)
    per_layer = t2 / 2.0                                   # upper bound: the probe's fixed cost (heads, embedding) is folded in
    B, L = cfg["bsz"], cfg["adapter_layer"]
    b = B
    while b > 1 and n_steps * b * L * per_layer > budget_s:
        b //= 2
    return b, per_layer


def cpu_baseline(cfg, budget_s=30.0):
    """Bounded CPU sample for the `cpu_baseline` object: ONE real full-depth fwd+bwd of the oracle at a bounded batch (no layer
    extrapolation), after a discarded 2-layer warm call."""
    threads = os.cpu_count() or 1
    B, L = cfg["bsz"], cfg["adapter_layer"]
    Bs, per_layer = _bounded_batch(cfg, threads, 1, budget_s)
    t = CpuOracleStep(cfg, L, threads).step(Bs)
    return dict(value=Bs / t, unit="samples/s", cores=threads, kind="port", extrapolated=False,
                sample=f"oracle fp32 fwd+bwd (CPU port of llama/model.py step), {WORKLOAD_NAMES.get(cfg['name'], cfg['name'])} shapes, ALL {L} layers, "
                       f"one step at batch {Bs} of {B} in {t:.1f} s (2-layer probe: {per_layer:.2f} s per layer per sample)")


def run_reference_arm(a, guard):
    """`--impl reference`: the reference's step is pure PyTorch with no native path and /root/reference does not exist on the GPU
    box; it is represented by the oracle port on the host cores. Every timed step is a REAL full-depth forward + backward (all
    layers, three objective streams, heads) at a bounded batch chosen so that K + W steps end within minutes; value = that
    batch / measured step time (nothing is extrapolated over layers). Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CONFIGS[a.config], name=a.config)
    threads = os.cpu_count() or 1
    B, L = cfg["bsz"], cfg["adapter_layer"]
    Bs, per_layer = _bounded_batch(cfg, threads, a.steps + a.warmup, 240.0)
    model = CpuOracleStep(cfg, L, threads)
    for i in range(a.warmup):
        model.step(Bs, seed=i)
    times = [model.step(Bs, seed=100 + i) for i in range(a.steps)]
    t = sum(times) / len(times)
    value = Bs / t
    line = {"impl": "reference", "metric": "7B NExT-QA train samples/s" if a.config == "7b-nextqa" else f"{a.config} train samples/s",
            "value": value, "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "extrapolated": False,
            "config": {"workload": WORKLOAD_NAMES[a.config], "batch_per_step": Bs, "config_batch": B, "layers": L,
                       "note": "full depth, bounded batch: samples/s = batch_per_step / measured step time"},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "extrapolated": False,
                             "sample": f"oracle (CPU port of llama/model.py step) fp32 fwd+bwd, all {L} layers, batch {Bs} of {B} per timed step; "
                                       f"mean of {a.steps} steps = {t:.2f} s (min {min(times):.2f}, max {max(times):.2f})"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    guard.emit(json.dumps(line))


def alt_build_leg(a):
    """The same resident-batch measurement with the OTHER operand build of the library (FVQA_DTYPE: bf16 when this process loaded
    fp16 and vice versa), in a child process on the same GPU right after this one's legs: the operand format is a property of
    the build, north_star names bf16, the default build is fp16 because that is the one that meets the gradient
    tolerance at full depth. Not the headline; outside every timed region of this process."""
    import subprocess
    from flipped_vqa_b200 import _lib
    other = "bf16" if _lib.DTYPE_NAME == "fp16" else "fp16"
    cmd = [sys.executable, os.path.abspath(__file__), "--config", a.config, "--steps", str(a.steps), "--warmup", str(a.warmup),
           "--no-e2e", "--no-padfree", "--no-cpu-baseline", "--no-eager-baseline", "--no-alt-build"]
    try:
        torch.cuda.empty_cache()
        r = subprocess.run(cmd, env={**os.environ, "FVQA_DTYPE": other}, capture_output=True, text=True, timeout=240)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        d = json.loads(lines[-1])
        return {"dtype": d["dtype"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                "tensor_util": d["tensor_util"]["value"], "gemm_tflops_in_step": d["roofline"]["achieved"], "gemm_frac_of_sustained_peak": d["roofline"]["frac"],
                "clocks": d["clocks"],
                "what": f"python bench.py with FVQA_DTYPE={other} (lib{'fvqa_bf16' if other == 'bf16' else 'fvqa'}.so), same GPU, same config / steps / warm-up, "
                        "resident batch; gradient parity of this build at full depth: DESIGN.md 5"}
    except Exception as e:          # informative leg: never lose the headline numbers over it
        return {"dtype": other, "value": None, "what": f"failed: {e}"}


def gpu_eager_baseline(model, cfg, dev, steps=3):
    """SURVEY 8(d) last row: the reference's op sequence (`llama/model.py:250-365`: three streams one after another, unfused
    attention with materialised scores, full-vocabulary logits, autograd) in stock eager PyTorch on the SAME GPU, in the
    library's operand dtype (fp16 = the reference's own), on the SAME weights as the product model (the oracle's state dict
    aliases the model's parameters) + fused AdamW. Outside every timed region of the product; N = 1 only."""
    from oracle import llama_vqa_oracle as O
    from flipped_vqa_b200 import _lib
    from flipped_vqa_b200.synthetic import synthetic_batch
    sd = {n: p.detach() for n, p in model.state_dict().items()}
    st = O.prepare_state(sd, frozen_dtype=_lib.H16, device=dev)
    params = SimpleNamespace(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"], norm_eps=1e-6,
                             max_seq_len=cfg["seqlen"], adapter_len=ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    opt = torch.optim.AdamW([st[n] for n in O.trainable_names(st)], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    batches = [synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=MAX_FEATS, seed=2000 + i) for i in range(2)]

    def step(i):
        losses = O.forward_losses(st, params, batches[i % 2], max_feats=MAX_FEATS, tau=TAU)
        sum(losses).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return losses

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": cfg["bsz"] / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "dtype": _lib.DTYPE_NAME, "steps": steps,
            "what": "oracle = the reference's op sequence under autograd in eager PyTorch (cuBLAS / ATen kernels) + fused AdamW, same GPU, same "
                    "weights and shapes, resident batch; not part of the product path"}


# ---------------------------------------------------------------------------------------------------
class _StdoutGuard:
    """Everything any library prints to fd 1 while the bench runs (e.g. NCCL's version banner) goes to stderr;
    `emit()` writes the ONE JSON line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    guard = _StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="7b-nextqa", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--no-padfree", action="store_true", help="skip the extra padding-free leg")
    ap.add_argument("--sample-layers", type=int, default=2, help="layers whose GEMM launches are event-timed inside the timed region")
    ap.add_argument("--chunk-layers", type=int, default=8, help="N > 1: layers per early adapter-gradient all-reduce message (dp.GradSync)")
    ap.add_argument("--chunk-ctas", type=int, default=0, help="N > 1: CTA cap of the NCCL communicator that carries the overlapped chunk messages (0 = one communicator)")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the eager-PyTorch-on-the-same-GPU comparator leg (N = 1)")
    ap.add_argument("--no-alt-build", action="store_true", help="skip the leg that runs the OTHER operand build (bf16 when this is fp16) on the same GPU (N = 1)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference_arm(a, guard)
        return

    import torch.distributed as dist
    from flipped_vqa_b200 import _lib, ops
    from flipped_vqa_b200.dp import DataParallel
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    from flipped_vqa_b200.synthetic import synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(CONFIGS[a.config], name=a.config)
    B, S = cfg["bsz"], cfg["seqlen"]

    torch.manual_seed(0)                                       # same frozen base on every rank
    params = ModelArgs(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"],
                       multiple_of=cfg["multiple_of"], norm_eps=1e-6, max_batch_size=32, max_seq_len=S,
                       adapter_len=ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    model = Transformer(params, make_args(), tokenizer=SyntheticTokenizer(cfg["vocab_size"]), device=dev)
    with torch.no_grad():                                      # zero-init gate1 would make every adapter gradient 0
        for blk in model.layers:
            blk.attention.gate1.normal_(0, 0.5)
    model.repack()
    net = DataParallel(model, chunk_layers=a.chunk_layers, chunk_ctas=a.chunk_ctas) if world > 1 else model
    trainables = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(trainables, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)

    batches = [synthetic_batch(B, S, cfg["vocab_size"], max_feats=MAX_FEATS, seed=1000 * rank + i) for i in range(4)]
    plans = [model.plan_batch(b) for b in batches]
    torch.cuda.synchronize()
    n_lab = sum(p.ce_total for p in plans) / len(plans)
    n_live = sum(p.n_live for p in plans) / len(plans)
    step_flops = flops_per_step(cfg, n_lab, n_live if model._engine.prune_last_layer else None)

    def step_resident(i):
        vqa, vaq, qav = net.forward_plan(plans[i % len(plans)])
        (vqa + vaq + qav).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    def step_e2e(i):
        vqa, vaq, qav = net(batches[i % len(batches)])
        loss = vqa + vaq + qav
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item()                                      # D2H read of the step's result

    step_ms = []        # per-step device times of the last timed() call on this rank (diagnostic: one-off stalls show up here)
    rank_ms = []        # device time of the timed region on every rank (N > 1)

    def dp_check():
        """N > 1, un-timed: hardware proof that the data-parallel gradient is the rank MEAN (DDP semantics, `train.py:115-117`).
        The same batch plan goes through backward twice on every rank - once with the exchange switched off (local gradient),
        once through DataParallel (chunked NCCL all-reduces overlapped with backward) - then every rank all-gathers the local
        flat buffers and compares its reduced buffer with their mean."""
        net._attach()
        sync = model.grad_sync
        model.grad_sync = None
        vqa, vaq, qav = model.forward_plan(plans[0])
        (vqa + vaq + qav).backward()
        local = model._grad_buffers.flat.clone()
        model.grad_sync = sync
        opt.zero_grad(set_to_none=True)
        vqa, vaq, qav = net.forward_plan(plans[0])
        (vqa + vaq + qav).backward()
        reduced = model._grad_buffers.flat.clone()
        opt.zero_grad(set_to_none=True)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean = torch.stack(gathered).double().mean(0)
        err = float((reduced.double() - mean).abs().max() / mean.abs().max())
        distinct = float((gathered[0] - gathered[-1]).abs().max()) > 0          # ranks really hold different gradients
        t = torch.tensor([err], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
        return {"status": "ok" if (worst < 1e-5 and distinct) else "FAILED", "max_rel_err_over_ranks": worst, "ranks_hold_distinct_gradients": distinct,
                "elements": int(local.numel()), "what": "post-all-reduce flat gradient vs mean of the all-gathered pre-reduce buffers"}

    def timed(fn, steps, sample_gemm=False):
        gc.collect()
        gc.disable()                    # a generational GC pause on one rank stalls every rank at the next all-reduce
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        if sample_gemm:
            ops.GEMM_TIMER = ops.GemmTimer()
        launches0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e0.record()
        for i in range(steps):
            fn(i)
            marks[i].record()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        gc.enable()
        ms = e0.elapsed_time(e1)
        step_ms.clear()
        step_ms.extend(round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks))
        launches = ops.LAUNCHES - launches0
        clocks = sampler.stop()
        timer, ops.GEMM_TIMER = ops.GEMM_TIMER, None
        rank_ms.clear()
        if world > 1:
            every = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(every, torch.tensor([ms], device=dev))
            rank_ms.extend(round(float(x.item()), 3) for x in every)
            ms = max(rank_ms)
        return ms, launches, clocks, timer

    L = cfg["adapter_layer"]
    chunk = a.chunk_layers if world > 1 else model._engine.adapter_grad_chunk
    quiet_layers = {0, L // 2} if a.sample_layers >= 2 else {L // 2}            # not the pruned last layer
    # N > 1: the all-reduce of the adapter rows of layers [c, c + chunk) is issued after layer c's backward and runs on NCCL's
    # stream while layer c - 1 is computed: sample such layers too, so SM contention from NCCL shows up in the GEMM numbers
    overlap_layers = {c - 1 for c in range(chunk, L, chunk)} if world > 1 else set()
    overlap_layers = set(sorted(overlap_layers)[-2:])
    model._engine.sample_layers = tuple(sorted(quiet_layers | overlap_layers)) if a.sample_layers > 0 else ()
    ClockSampler._nvml(local)                                   # NVML set-up before the warm-up, not between warm-up and timing
    dpc = dp_check() if world > 1 else None
    for i in range(a.warmup):
        step_resident(i)
    if world > 1:
        model.grad_sync.timing = []
        msg0 = model.grad_sync.messages
    ms, launches, clocks, gemm_timer = timed(step_resident, a.steps, sample_gemm=a.sample_layers > 0)
    gemm = gemm_timer.summary(quiet_layers) if gemm_timer is not None else None
    gemm_overlap = gemm_timer.summary(overlap_layers) if (gemm_timer is not None and overlap_layers) else None
    ms_per_step = ms / a.steps
    resident_step_ms = list(step_ms)
    resident_rank_ms = list(rank_ms)
    value = world * B / (ms_per_step * 1e-3)
    comm = None
    if world > 1:
        waits = [e0.elapsed_time(e1) for e0, e1 in model.grad_sync.timing]
        model.grad_sync.timing = None
        mine = torch.tensor([sum(waits) / max(len(waits), 1), max(waits) if waits else 0.0], device=dev)
        every = [torch.zeros(2, device=dev) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank_step = [x / a.steps for x in resident_rank_ms]
        comm = {"comm_exposed_ms_per_step_by_rank": [round(float(x[0]), 4) for x in every],
                "comm_exposed_ms_worst_step_by_rank": [round(float(x[1]), 4) for x in every],
                "step_ms_by_rank": [round(x, 3) for x in per_rank_step], "rank_skew_ms_per_step": round(max(per_rank_step) - min(per_rank_step), 4),
                "messages_per_step": (model.grad_sync.messages - msg0) / a.steps,
                "what": "comm_exposed = CUDA-event time the compute stream spends in GradSync.finish() waiting for NCCL (late message + rank "
                        "skew: a rank that finishes backward early waits here for the slowest one); step_ms_by_rank = each rank's own device "
                        "time of the timed region / steps"}

    if a.no_e2e:
        ms_e, clocks_e = float("nan"), None
    else:
        for i in range(2):
            step_e2e(i)
        ms_e, _, clocks_e, _ = timed(step_e2e, a.steps)
    e2e_value = world * B / (ms_e / a.steps * 1e-3)

    # Host side: wall-clock time to ENQUEUE one resident step with an empty GPU queue (no sync inside; N = 1 only: at N > 1 the
    # ranks would have to stay in lock-step for the NCCL calls). Outside every timed region.
    host = None
    if world == 1:
        import time as _time
        ts = []
        for i in range(5):
            torch.cuda.synchronize()
            t0 = _time.perf_counter()
            step_resident(i)
            ts.append((_time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        host = {"enqueue_ms_per_step": round(sorted(ts)[len(ts) // 2], 2), "launches_per_step": launches / a.steps,
                "what": "median host wall-clock time of forward_plan + backward + AdamW with an empty GPU queue and no synchronisation"}

    # HBM-bound kernels (north_star: achieved GB/s of the norm / attention kernels against the HBM roofline): CUDA events around
    # every C-ABI call of three extra resident steps (outside every timed region; N = 1). Algorithmic bytes per launch;
    # event time includes the launch gap in front of the kernel, so these are lower bounds of the kernels' own rates.
    hbm_kernels = None
    if world == 1 and not a.no_e2e:
        ops.OP_TIMER = ops.OpTimer()
        for i in range(3):
            step_resident(i)
        torch.cuda.synchronize()
        agg, ops.OP_TIMER = ops.OP_TIMER.summary(), None
        T_all, dm = 3 * B * S, cfg["dim"]
        per_launch = {"rmsnorm_fwd": T_all * dm * 6.0, "rmsnorm_bwd": T_all * dm * 16.0, "attn_fwd": T_all * dm * 8.0, "attn_bwd": T_all * dm * 16.0}
        hbm_peak = measured_peaks()["hbm_gbs"]
        hbm_kernels = {"peak_gbs": hbm_peak, "what": "algorithmic bytes per launch / in-step CUDA-event time per launch (events include the launch gap); "
                                                     "rmsnorm fwd 6d, bwd 16d, attention fwd 8d, bwd 16d bytes per token"}
        for name, nbytes in per_launch.items():
            if name in agg and agg[name][0] > 0:
                us = agg[name][1] / agg[name][0] * 1e3
                # the pruned last layer runs its second norm on a handful of rows: the per-launch average is slightly optimistic in bytes
                hbm_kernels[name] = {"avg_us": round(us, 1), "gbs": round(nbytes / (us * 1e-6) / 1e9, 0), "frac": round(nbytes / (us * 1e-6) / 1e9 / hbm_peak, 3),
                                     "launches_per_step": agg[name][0] / 3}

    # Extra leg (not the headline): the opt-in padding-free row set (StepEngine.skip_pad_rows) on the same resident batches
    padfree = None
    if not a.no_padfree:
        model._engine.skip_pad_rows = True
        sampled_layers, model._engine.sample_layers = model._engine.sample_layers, ()
        plans_dense, plans = plans, [model.plan_batch(b) for b in batches]      # with the padding-free row maps
        torch.cuda.synchronize()
        for i in range(max(a.warmup, 3)):
            step_resident(i)
        ms_p, _, _, _ = timed(step_resident, a.steps)
        model._engine.skip_pad_rows = False
        model._engine.sample_layers = sampled_layers
        rows_p = sum(p.T_c for p in plans) / len(plans)
        n_live_p = sum(p.n_live_k for p in plans) / len(plans)
        plans = plans_dense
        fl_p = flops_per_step(cfg, n_lab, n_live_p if model._engine.prune_last_layer else None, rows=rows_p)
        padfree = {"value": world * B / (ms_p / a.steps * 1e-3), "unit": "samples/s", "ms_per_step": ms_p / a.steps,
                   "rows_per_step": rows_p, "rows_dense": 3 * B * S, "flops_per_step": fl_p,
                   "note": "opt-in StepEngine.skip_pad_rows: norms / frozen GEMMs / SwiGLU skip the rows after each sequence's last "
                           "loss-relevant position (same losses and gradients); FLOPs counted on the rows executed"}

    peaks = measured_peaks()
    if padfree is not None:
        padfree["tensor_util"] = padfree["flops_per_step"] / (padfree["ms_per_step"] * 1e-3) / 1e12 / peaks["bf16_tflops"]
    if rank == 0:
        line = {
            "metric": "7B NExT-QA train samples/s" if a.config == "7b-nextqa" else f"{a.config} train samples/s",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "host": host, "hbm_kernels": hbm_kernels,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": _lib.DTYPE_NAME, "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[a.config], "operands": f"{_lib.DTYPE_NAME} weights / activations / gradients (tcgen05 kind::f16, fp32 accumulate), fp32 residual stream and trainables", "global_batch": world * B, "seq_len": S, "parallelism": f"dp{world}", "allreduce_chunk_layers": (a.chunk_layers if world > 1 else None), "allreduce_chunk_ctas": (a.chunk_ctas if world > 1 else None), "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")},
                       "l2": "working set (13.5 GB frozen weights read in both passes + 2.1 GB transposed Wk|Wv + 9 GB saved activations per step at 7B) >> 126 MB L2; no explicit flush",
                       "objectives": "vqa+vaq+qav", "optimizer": "AdamW(fused) on 4.5M trainables", "flops_per_step": step_flops,
                       "labelled_rows_per_step": n_lab,
                       "last_layer_live_rows": (n_live if model._engine.prune_last_layer else None),
                       "pad_rows": "computed (the reference's row set: every position of every sequence)"},
            "tensor_util": {"value": step_flops / (ms_per_step * 1e-3) / 1e12 / peaks["bf16_tflops"],
                            "achieved_tflops": step_flops / (ms_per_step * 1e-3) / 1e12, "peak_tflops": peaks["bf16_tflops"],
                            "peak_sustained_tflops": peaks["bf16_tflops_sustained"], "peak_source": peaks["source"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": plans[0].h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e / a.steps, "clocks": clocks_e},
            "gpu_launches": launches,
            "step_ms_rank0": resident_step_ms,
            "padding_free": padfree,
        }
        if gemm is not None:
            tr = gemm_traffic() if a.config == "7b-nextqa" else None
            line["roofline"] = {"bound": "tensor", "kernel": "gemm_nt_pair_kernel (tcgen05.mma.cta_group::2, TMA, TMEM; 2x2-cluster TMA-multicast variant on the N = 4096 GEMMs)", "achieved": gemm["tflops"],
                                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": gemm["tflops"] / peaks["bf16_tflops_sustained"],
                                "frac_of_burst_peak": gemm["tflops"] / peaks["bf16_tflops"], "peak_kind": f"{peaks['source']} sustained (kernel timed inside a long step)",
                                "traffic": (tr["avg_dram_bytes_per_launch"] if tr else None),
                                "traffic_note": (tr["note"] if tr else "no ncu capture for this config"),
                                "launches_sampled": gemm["launches"], "avg_launch_ms": gemm["avg_ms"],
                                "avg_flops_per_launch": gemm["avg_flops"], "sampled_layers": sorted(quiet_layers),
                                "gemm_share_of_step": gemm["total_ms"] / len(quiet_layers) * L / (ms_per_step * a.steps)}
            if gemm_overlap is not None:
                line["roofline"]["overlapped_with_allreduce"] = {"sampled_layers": sorted(overlap_layers), "achieved": gemm_overlap["tflops"],
                                                                 "avg_launch_ms": gemm_overlap["avg_ms"], "launches_sampled": gemm_overlap["launches"],
                                                                 "vs_quiet_layers": gemm_overlap["tflops"] / gemm["tflops"]}
        if dpc is not None:
            line["dp_check"] = dpc
        if comm is not None:
            line["comm"] = comm
        if world == 1 and not a.no_eager_baseline:
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(model, cfg, dev)
                line["gpu_eager_baseline"]["speedup_of_this_repo"] = value / line["gpu_eager_baseline"]["value"]
            except Exception as e:
                line["gpu_eager_baseline"] = {"value": None, "unit": "samples/s", "what": f"failed: {e}"}
        if world == 1 and not a.no_alt_build and not a.no_e2e:      # (--no-e2e = profiling runs: no child processes under ncu)
            line["alt_build"] = alt_build_leg(a)
        if world == 1 and not a.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(cfg)
            except Exception as e:  # the baseline is informative; never lose the GPU numbers over it
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
