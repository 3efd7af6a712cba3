#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric : decode tokens/sec, LLaMA-3-8B Q4_K_M, single stream, ctx = 2048 (configs[1]) on N B200s of one node.
step   : one burst of BURST greedy tokens at kv positions [ctx-BURST-1, ctx-1) of a synthetic LLaMA-3-8B-shaped
         Q4_K_M GGUF (random quantized blocks in the reference's tensor-type mixture — no model file exists
         offline). The KV cache is filled for all earlier positions by a real (untimed) decode pass.
value  : whole-job tokens/s with inputs resident in HBM: device-resident greedy loop (CUDA-graph replay per
         token, arg-max on device), timed with CUDA events on the engine's stream.
e2e    : same metric through the C-ABI with HOST buffers: per token b200_decode(token id from host) -> logits
         to host (pinned D2H inside the call) -> host arg-max; wall clock around the calls.
roofline: dominant kernel = k_matvec<gate/up> (47 % of the bytes of a token); achieved = algorithmic bytes of one
         launch / its mean device time, measured live with a CUDA-event pair around every launch of an
         un-graphed token (b200_profile_token); peak = MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the reference's own CPU llama_decode (oracle/_ref, unmodified sources) on the
         host cores, same GGUF, on a bounded sample (decode at n_kv ~ 64..).

N > 1 (torchrun, one rank per GPU): layers are split across ranks (stage r = layers [r*L/N, (r+1)*L/N)); one
ncclSend/ncclRecv of the residual stream per boundary per token plus the 4-byte token hand-back; single stream,
so stages run one after the other — value is NOT expected to rise with N, "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from booster_b200 import gguf_io as G  # noqa: E402

CTX = 2048
BURST = 64
CFG_NAME, FTYPE = "llama3-8b", "Q4_K_M"
MODEL_DIR = os.environ.get("B200_TMP", os.path.join(tempfile.gettempdir(), "b200_models"))


def model_path():
    os.makedirs(MODEL_DIR, exist_ok=True)
    p = os.path.join(MODEL_DIR, f"{CFG_NAME}_{FTYPE}_s1234.gguf")
    if not os.path.exists(p):
        tmp = p + f".tmp{os.getpid()}"
        G.synth_llama(tmp, G.CONFIGS[CFG_NAME], FTYPE, seed=1234, source="blocks")
        os.replace(tmp, p)
    return p


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.p, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(path, steps, warmup, sample_tokens=None, n_prompt=64):
    """the reference's own CPU llama_decode on the host cores; returns (tokens/s, info). A step is `sample_tokens` greedy
    decode tokens (default: up to 32, capped so that the whole run stays inside ctx 512 — BASELINE.md §2: 64-token prompt,
    then >= 128 decode steps when the step count allows)"""
    if sample_tokens is None:
        sample_tokens = max(1, min(32, 400 // max(1, steps)))
    from oracle import ref
    cores = os.cpu_count() or 1
    best = None
    # "threads = host cores" (BASELINE.json configs[0]); OpenMP barriers can make fewer threads faster on big hosts,
    # so give the reference its best of {cores, cores/2 (physical), 32} — all the host threads it can use well
    cands = sorted({cores, max(1, cores // 2), min(cores, 32)}, reverse=True)
    prompt = np.random.default_rng(42).integers(0, G.CONFIGS[CFG_NAME].n_vocab, size=n_prompt).tolist()
    for th in cands:
        r = ref.RefModel(path, n_ctx=512, n_batch=512, n_threads=th)
        r.decode(prompt, 0)                        # prefill (untimed; also pages the file in)
        pos = n_prompt
        tok = 1
        for _ in range(max(1, warmup)):
            r.decode([tok], pos); pos += 1
        t0 = time.perf_counter()
        for _ in range(2):
            r.decode([tok], pos); pos += 1
        quick = 2 / (time.perf_counter() - t0)
        r.close()
        if best is None or quick > best[1]:
            best = (th, quick)
    th = best[0]
    r = ref.RefModel(path, n_ctx=512, n_batch=512, n_threads=th)
    r.decode(prompt, 0)
    pos = n_prompt
    for _ in range(max(1, warmup)):
        r.decode([1], pos); pos += 1
    r.reset_timings()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        for _ in range(sample_tokens):
            r.decode([1], pos); pos += 1; n += 1
    dt = time.perf_counter() - t0
    tm = r.timings()
    r.close()
    lib_tps = 1e3 * tm["n_eval"] / tm["t_eval_ms"] if tm["t_eval_ms"] > 0 else None
    return n / dt, {"cores": th, "host_cores": cores, "kind": "reference", "variant": ref.variant(), "tokens_per_step": sample_tokens,
                    "sample": f"{n} greedy decode steps (batch 1, n_kv {n_prompt}..{pos}) after a {n_prompt}-token prefill, ctx 512, "
                              f"threads={th} (best of {cands})", "llama_timings_tok_s": lib_tps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (A/B runs of the GPU path only)")
    ap.add_argument("--value-only", action="store_true", help="A/B runs: device-resident value only (no e2e / roofline / cpu legs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = G.CONFIGS[CFG_NAME]
    config = {"workload": f"LLaMA-3-8B-shaped synthetic GGUF, {FTYPE}, single-stream greedy decode, ctx={CTX}, "
                          f"burst of {BURST} tokens at n_kv {CTX - BURST}..{CTX - 1}",
              "n_layer": cfg.n_layer, "n_embd": cfg.n_embd, "n_vocab": cfg.n_vocab, "ctx": CTX, "burst": BURST,
              "parallelism": f"layer-split pp{world}" if world > 1 else "single GPU",
              "l2": "inputs larger than L2 (4.6 GB of weights streamed per token vs 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        path = model_path()
        tps, info = cpu_reference_run(path, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "decode tokens/sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["tokens_per_step"] / tps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "q4_K x q8_K int8 dot, f32 accumulate", "data": "synthetic",
                "config": config, "cpu_baseline": dict(info, value=tps, unit="tokens/s"),
                "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from booster_b200 import engine, pipeline

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    path = model_path() if rank == 0 else None
    if world > 1:
        dist.barrier()
        path = model_path()
    # stage = contiguous layer range (cpp/src/llama.cpp:5932-5968 with equal proportions)
    lb, le = pipeline.stage_range(cfg.n_layer, rank, world)
    m = engine.Model(path, device=local_rank, layer_begin=lb, layer_end=le)
    c = engine.Context(m, CTX)
    if world > 1:
        c.comm_init(rank, world, pipeline.share_unique_id(dist, engine.comm_unique_id))
    gen = (lambda tok, pos, n: c.pipeline_generate_greedy(tok, pos, n)) if world > 1 else (lambda tok, pos, n: c.generate_greedy(tok, pos, n))

    # fill the KV cache with a real decode pass up to the burst start (untimed)
    pos0 = CTX - BURST - 1
    filled = 0
    tok = 1
    while filled < pos0:
        n = min(512, pos0 - filled)
        out = gen(tok, filled, n)
        tok = int(out[-1]); filled += n
    for _ in range(args.warmup):
        gen(tok, pos0, BURST)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    clocks = ClockSampler(local_rank)
    l0 = c.kernel_launches()
    sync_all()
    clocks.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    ids = None
    for _ in range(args.steps):
        ids = gen(tok, pos0, BURST)
        dev_ms += c.last_device_ms()
    sync_all()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    launches = c.kernel_launches() - l0
    # device time from CUDA events on the engine's stream (every rank's span covers the whole burst: a stage's stream
    # sits in ncclRecv while the other stages work), max over ranks; the barrier-bracketed wall clock is reported beside it
    elapsed = dev_ms / 1e3
    if world > 1:
        elapsed = pipeline.max_over_ranks(dist, elapsed, device="cuda")
        launches = pipeline.sum_over_ranks(dist, launches, device="cuda")
    tokens = args.steps * BURST
    tps = tokens / elapsed

    line = {"metric": "decode tokens/sec", "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "q4_K/q6_K x q8_K int8 dot (dp4a), f32 accumulate, f16 KV", "data": "synthetic",
            "config": config, "gpu_launches": launches, "clocks": clk,
            "wall_clock_tokens_per_s": tokens / wall,
            # the arithmetic is bit-exact and the inputs are fixed: the burst's token ids are the same at every N
            "burst_ids_crc32": zlib.crc32(np.ascontiguousarray(ids, dtype=np.int32).tobytes())}

    if rank == 0 and world == 1 and not args.value_only:
        peak, peak_src = peaks()
        # ---- end to end through the C-ABI with host buffers
        n_e2e = min(args.steps, 4) * BURST
        lg = c.decode([tok], pos0)
        t0 = time.perf_counter()
        p = pos0
        for i in range(n_e2e):
            t = int(np.argmax(lg))
            p = pos0 + (i % BURST)
            lg = c.decode([t], p)
        e2e_dt = time.perf_counter() - t0
        line["e2e"] = {"value": n_e2e / e2e_dt, "unit": "tokens/s",
                       "h2d_bytes_per_step": 16 * BURST, "d2h_bytes_per_step": 4 * cfg.n_vocab * BURST,
                       "what": "b200_decode(token from host) -> logits to host -> host arg-max, per token"}
        # ---- live per-kernel roofline (event pair around every launch of an un-graphed token at n_kv ~ ctx)
        acc = {}
        reps = 4
        for i in range(reps + 1):
            prof = c.profile_token(tok, pos0 + BURST // 2)
            if i == 0:
                continue   # warm-up
            for k, (ms, n) in prof.items():
                a = acc.setdefault(k, [0.0, 0]); a[0] += ms; a[1] += n
        types = G.tensor_types(cfg, FTYPE)
        E, FF = cfg.n_embd, cfg.n_ff
        gu_bytes = np.mean([G.row_bytes(types[f"blk.{i}.ffn_gate.weight"], E) * FF + G.row_bytes(types[f"blk.{i}.ffn_up.weight"], E) * FF
                            for i in range(cfg.n_layer)]) + 4 * E
        gu_ms = acc["gate_up"][0] / acc["gate_up"][1]
        achieved = gu_bytes / (gu_ms * 1e-3) / 1e9
        n_kv_mean = pos0 + BURST / 2
        bytes_tok = m.weight_bytes + G.kv_bytes_per_token(cfg, int(n_kv_mean))
        tot_ms = sum(v[0] for v in acc.values()) / reps
        # DRAM traffic of one gate|up launch from the committed ncu --set full capture (profiles/traffic.json, written by
        # scripts/ncu_summary.py): dram__bytes_read.sum + dram__bytes_write.sum
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)["gate_up"]
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "kernel": "k_matvec<EPI_SILU> (ffn gate|up, fused RMSNorm + Q8_K quant + SiLU*mul)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "peak_source": peak_src, "bytes_per_launch": float(gu_bytes), "ms_per_launch": gu_ms,
                            "share_of_token_time": acc["gate_up"][0] / reps / tot_ms}
        # the same kernel outside the token's dependency chain: the gate|up launches of all 32 layers back to back
        # (2.1 GB of distinct tiles), one event pair around the lot
        iso_ms, iso_n = c.profile_kind("gate_up", pos0 + BURST // 2, reps=8)
        line["roofline"]["isolated"] = {"ms_per_launch": iso_ms, "launches_timed": iso_n, "achieved": gu_bytes / (iso_ms * 1e-3) / 1e9,
                                        "frac": gu_bytes / (iso_ms * 1e-3) / 1e9 / peak,
                                        "what": "gate|up launches of every layer back to back (distinct weights per launch, PDL), one CUDA-event pair"}
        line["token_roofline"] = {"bytes_per_token": int(bytes_tok), "achieved_gbs": bytes_tok * tps / 1e9,
                                  "frac_of_peak": bytes_tok * tps / 1e9 / peak, "roofline_tokens_per_s": peak * 1e9 / bytes_tok}
        line["kernel_ms_per_token"] = {k: round(v[0] / reps, 4) for k, v in acc.items()}
        # ---- the reference's CPU llama_decode on this host, bounded sample
        try:
            if args.no_cpu:
                raise RuntimeError("skipped (--no-cpu)")
            cpu_tps, info = cpu_reference_run(path, steps=4, warmup=1)
            line["cpu_baseline"] = dict(info, value=cpu_tps, unit="tokens/s")
        except Exception as e:  # the oracle library failing to load must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    c.close(); m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
