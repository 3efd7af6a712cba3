#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

metric : decode tokens/sec, single stream. `--config` picks the workload; the default is the configuration the metric is
         quoted on (configs[1]: LLaMA-3-8B Q4_K_M, ctx 2048, 1 x B200). The others are the remaining GPU configs of
         BASELINE.json (configs[2..4]); configs[0] is the reference's own CPU case and is what `--impl reference` runs.

           8b-q4km-2048        LLaMA-3-8B Q4_K_M, single-stream decode at the end of ctx 2048            (default)
           8b-q8_0-prefill512  LLaMA-3-8B Q8_0: one 512-token prompt batch, then 1024 generated tokens
           mistral-q5km-8192   Mistral-7B Q5_K_M, single-stream decode at the end of ctx 8192
           70b-q4km-4096       LLaMA-3-70B Q4_K_M, ctx 4096 — one GPU (it fits) or layer-split with --gpus 8

step   : one burst of `burst` greedy tokens at the last kv positions of the context of a synthetic GGUF with the
         model's exact shapes (random quantized blocks in the reference's tensor-type mixture — no model file exists
         offline). The KV cache is filled for all earlier positions by a real (untimed) decode pass.
         (8b-q8_0-prefill512: a step is the 512-token prompt + 1024 generated tokens; value = generated tokens/s,
         the prompt is reported beside it.)
value  : whole-job tokens/s with inputs resident in HBM: device-resident greedy loop (CUDA-graph replay per token,
         arg-max on device), timed with CUDA events on the engine's stream.
e2e    : the same metric through the reference-facing boundary with HOST buffers: init -> initContext -> doInference ->
         status (the nine symbols of include/bridge.h) with the prompt as host text; generated tokens are counted by a
         status() poller, the way the Go server streams them, and timed from the first generated token to the last
         (every token: host state -> device, sampled id -> host). The additive b200_decode figure (host token in, host
         logits out, host arg-max) is reported beside it.
roofline: dominant kernel = the ffn gate|up mat-vec (47 % of the bytes of an 8B Q4_K_M token); achieved = algorithmic
         bytes of one launch / its mean device time, measured live with a CUDA-event pair around every launch of an
         un-graphed token (b200_profile_token); peak = MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the reference's own CPU llama_decode (oracle/_ref, unmodified sources) on the
         host cores, same 8B GGUF, on a bounded sample (BASELINE.json configs[0]: ctx 512).

N > 1 (torchrun, one rank per GPU): layers are split across ranks (stage r = layers [r*L/N, (r+1)*L/N)); one hand-off
of the residual stream per boundary per token plus the 4-byte token hand-back; single stream, so stages run one after
the other — value is NOT expected to rise with N, "scaling": "strong".
"""
import argparse
import ctypes as C
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

MODEL_DIR = os.environ.get("B200_TMP", os.path.join(tempfile.gettempdir(), "b200_models"))

BENCH_CONFIGS = {
    "8b-q4km-2048": dict(model="llama3-8b", ftype="Q4_K_M", ctx=2048, burst=64, label="LLaMA-3-8B"),
    "8b-q8_0-prefill512": dict(model="llama3-8b", ftype="Q8_0", ctx=1536, prefill=512, burst=1024, label="LLaMA-3-8B"),
    "mistral-q5km-8192": dict(model="mistral-7b", ftype="Q5_K_M", ctx=8192, burst=64, label="Mistral-7B"),
    "70b-q4km-4096": dict(model="llama3-70b", ftype="Q4_K_M", ctx=4096, burst=32, label="LLaMA-3-70B", share_period=2),
}
DEFAULT_CONFIG = "8b-q4km-2048"
# module-level names kept for scripts that import bench (scripts/bridge_speed.py)
CTX, BURST, CFG_NAME, FTYPE = 2048, 64, "llama3-8b", "Q4_K_M"


def model_path(cfg_name=CFG_NAME, ftype=FTYPE, share_period=0):
    os.makedirs(MODEL_DIR, exist_ok=True)
    p = os.path.join(MODEL_DIR, f"{cfg_name}_{ftype}_s1234{'_sh%d' % share_period if share_period else ''}.gguf")
    if not os.path.exists(p):
        tmp = p + f".tmp{os.getpid()}"
        G.synth_llama(tmp, G.CONFIGS[cfg_name], ftype, seed=1234, source="blocks", share_period=share_period)
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
    pos_first = pos
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
                    "n_kv_first": pos_first, "n_kv_last": pos, "ctx": 512,
                    "sample": f"{n} greedy decode steps (batch 1, n_kv {pos_first}..{pos}) after a {n_prompt}-token prefill, ctx 512, "
                              f"threads={th} (best of {cands})", "llama_timings_tok_s": lib_tps}


def bridge_e2e(path, ctx, n_prompt, n_gen, reps):
    """decode tokens/s through the nine bridge symbols: doInference on a host text prompt of token ids, generated tokens
    counted by a status() poller (the Go server's streaming loop, pkg/server/server.go:842-863); the timed region runs
    from the first generated token's appearance to the last one's. Sampler arguments: Janus at scale = hi = lo = 1.0,
    i.e. the reference's deterministic setting."""
    from booster_b200 import _lib
    L = _lib.lib()
    L.init(b"", b"")
    h = L.initContext(7, path.encode(), 1, 0, 100, 0, 0, 0, ctx, n_gen, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
    if not h:
        raise RuntimeError("initContext failed")
    rng = np.random.default_rng(4242)
    best = None
    for rep in range(reps + 1):                   # rep 0 is the warm-up (graph capture, first touch): a short prompt
        prompt = " ".join(str(int(t)) for t in rng.integers(0, 1000, size=n_prompt if rep else 16)).encode()
        job = f"bench-e2e-{rep}".encode()
        marks = {}
        done = threading.Event()

        def poll():
            # count the pieces published so far: each piece of a no_vocab model is "<id> "
            while not done.is_set():
                n = L.status(job).count(b" ")
                now = time.perf_counter()
                if n > n_prompt and "first" not in marks:
                    marks["first"] = (now, n - n_prompt)
                if n > n_prompt:
                    marks["last"] = (now, n - n_prompt)
                # a streaming client polls at intervals (the Go server answers status requests, pkg/server/server.go:842-863);
                # a spinning poller starves doInference's own lock / this process's GIL hand-back by tens of ms at job end
                time.sleep(0.0002)

        th = threading.Thread(target=poll)
        th.start()
        ret = L.doInference(7, h, job, b"", prompt)
        t_end = time.perf_counter()
        done.set(); th.join()
        n_out = L.status(job).count(b" ") - n_prompt
        pu, gu = C.c_double(), C.c_double()
        L.b200_job_timing_us(job, C.byref(pu), C.byref(gu))
        if rep == 0 or "first" not in marks or n_out < 2:
            continue
        t_first, n_first = marks["first"]
        tps = (n_out - n_first) / (t_end - t_first)
        cand = {"value": tps, "generated": n_out, "returned": int(ret), "job_timing_us_per_token": gu.value,
                "prompt_us_per_token": pu.value}
        if best is None or cand["value"] > best["value"]:
            best = cand
    if best is None:
        raise RuntimeError("no generated tokens observed")
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(BENCH_CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (A/B runs of the GPU path only)")
    ap.add_argument("--value-only", action="store_true", help="A/B runs: device-resident value only (no e2e / roofline / cpu legs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    bc = BENCH_CONFIGS[args.config]
    cfg = G.CONFIGS[bc["model"]]
    ftype, ctx, burst, prefill = bc["ftype"], bc["ctx"], bc["burst"], bc.get("prefill", 0)
    data_note = ("synthetic (random quantized blocks in the reference's tensor-type mixture for %s; weight values do not "
                 "enter the timing)" % ftype)

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference's CPU path on BASELINE.json configs[0]: the 8B Q4_K_M file, ctx 512 — the config says what ran
        path = model_path()
        tps, info = cpu_reference_run(path, args.steps, args.warmup)
        c8 = G.CONFIGS[CFG_NAME]
        config = {"workload": f"LLaMA-3-8B-shaped synthetic GGUF, {FTYPE}, single-stream greedy decode on the host CPU, ctx=512, "
                              f"{info['tokens_per_step']} tokens per step at n_kv {info['n_kv_first']}..{info['n_kv_last']}, threads={info['cores']} "
                              f"(BASELINE.json configs[0]; the GPU arm's ctx-2048 positions would take the CPU minutes per step)",
                  "n_layer": c8.n_layer, "n_embd": c8.n_embd, "n_vocab": c8.n_vocab, "ctx": 512, "burst": info["tokens_per_step"],
                  "parallelism": f"{info['cores']} host threads (OpenMP)", "l2": "n/a (CPU)"}
        line = {"impl": "reference", "metric": "decode tokens/sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["tokens_per_step"] / tps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "q4_K x q8_K int8 dot, f32 accumulate", "data": data_note,
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
    sp = bc.get("share_period", 0)
    path = model_path(bc["model"], ftype, sp) if rank == 0 else None
    if world > 1:
        dist.barrier()
        path = model_path(bc["model"], ftype, sp)
    # stage = contiguous layer range (cpp/src/llama.cpp:5932-5968 with equal proportions)
    lb, le = pipeline.stage_range(cfg.n_layer, rank, world)
    m = engine.Model(path, device=local_rank, layer_begin=lb, layer_end=le)
    c = engine.Context(m, ctx)
    handoff = "single GPU"
    if world > 1:
        c.comm_init(rank, world, pipeline.share_unique_id(dist, engine.comm_unique_id))
        handoff = "ncclSend/ncclRecv per stage boundary"
        if os.environ.get("BOOSTER_B200_P2P", "1") != "0":
            # direct NVLink stores into the next stage's inbox (CUDA IPC) instead of NCCL point-to-point
            if pipeline.connect_peer_handoff(dist, c, rank, world):
                handoff = "peer stores over NVLink into the next stage's inbox (CUDA IPC) + sequence flag, per stage boundary"
    gen = (lambda tok, pos, n: c.pipeline_generate_greedy(tok, pos, n)) if world > 1 else (lambda tok, pos, n: c.generate_greedy(tok, pos, n))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    prompt = np.random.default_rng(42).integers(0, cfg.n_vocab, size=max(1, prefill)).tolist()
    prefill_ms = []
    if prefill:
        # config 3: the step is prompt batch + generation, both timed; positions restart at 0 every step
        pos0 = prefill
        workload = (f"{bc['label']}-shaped synthetic GGUF, {ftype}: one {prefill}-token prompt batch (b200_decode, batch > 1 arithmetic) "
                    f"then {burst} greedy tokens at n_kv {prefill}..{prefill + burst - 1}")

        def one_step():
            t0 = time.perf_counter()
            if world > 1:
                c.pipeline_decode(prompt, 0)
            else:
                c.decode(prompt, 0, want_logits=False)
            torch.cuda.synchronize()
            prefill_ms.append(1e3 * (time.perf_counter() - t0))
            return gen(1, pos0, burst)
    else:
        # fill the KV cache with a real decode pass up to the burst start (untimed)
        pos0 = ctx - burst - 1
        workload = (f"{bc['label']}-shaped synthetic GGUF, {ftype}, single-stream greedy decode, ctx={ctx}, "
                    f"burst of {burst} tokens at n_kv {ctx - burst}..{ctx - 1}")
        filled = 0
        tok = 1
        while filled < pos0:
            n = min(512, pos0 - filled)
            out = gen(tok, filled, n)
            tok = int(out[-1]); filled += n

        def one_step():
            return gen(tok, pos0, burst)

    config = {"workload": workload, "name": args.config, "n_layer": cfg.n_layer, "n_embd": cfg.n_embd, "n_vocab": cfg.n_vocab,
              "ctx": ctx, "burst": burst, "parallelism": f"layer-split pp{world}" if world > 1 else "single GPU", "hand_off": handoff,
              "l2": f"inputs larger than L2 ({m.weight_bytes / 1e9:.1f} GB of weights streamed per token on this rank vs 126 MB L2)"}
    for _ in range(args.warmup):
        one_step()
    prefill_ms.clear()

    clocks = ClockSampler(local_rank)
    l0 = c.kernel_launches()
    sync_all()
    clocks.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    ids = None
    for _ in range(args.steps):
        ids = one_step()
        dev_ms += c.last_device_ms()
    sync_all()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    launches = c.kernel_launches() - l0
    # device time from CUDA events on the engine's stream (every rank's span covers the whole burst: a stage's stream
    # waits for its predecessor while the other stages work), max over ranks; the barrier-bracketed wall clock is beside it
    elapsed = dev_ms / 1e3
    if world > 1:
        elapsed = pipeline.max_over_ranks(dist, elapsed, device="cuda")
        launches = pipeline.sum_over_ranks(dist, launches, device="cuda")
    tokens = args.steps * burst
    tps = tokens / elapsed

    dtypes = {"Q4_K_M": "q4_K/q6_K", "Q5_K_M": "q5_K/q6_K", "Q8_0": "q8_0"}
    line = {"metric": "decode tokens/sec", "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": f"{dtypes[ftype]} x {'q8_0' if ftype == 'Q8_0' else 'q8_K'} int8 dot (dp4a), f32 accumulate, f16 KV",
            "data": data_note, "config": config, "gpu_launches": launches, "clocks": clk,
            "wall_clock_tokens_per_s": (tokens + args.steps * prefill) / wall,
            # the arithmetic is bit-exact and the inputs are fixed: the burst's token ids are the same at every N
            "burst_ids_crc32": zlib.crc32(np.ascontiguousarray(ids, dtype=np.int32).tobytes())}
    if prefill:
        pm = float(np.mean(prefill_ms))
        line["prefill"] = {"tokens": prefill, "ms": pm, "tokens_per_s": 1e3 * prefill / pm,
                           "what": "b200_decode of the whole prompt batch, wall clock incl. synchronisation"}
        line["ms_per_step"] = 1e3 * elapsed / args.steps + pm
    n_kv_mean = pos0 + burst / 2
    peak, peak_src = peaks()
    if world > 1:
        # whole-job bytes: every rank's stage weights + the KV rows of all layers
        wb = pipeline.sum_over_ranks(dist, int(m.weight_bytes), device="cuda")
    else:
        wb = m.weight_bytes
    bytes_tok = wb + G.kv_bytes_per_token(cfg, int(n_kv_mean))
    line["token_roofline"] = {"bytes_per_token": int(bytes_tok), "achieved_gbs": bytes_tok * tps / 1e9,
                              "frac_of_peak": bytes_tok * tps / 1e9 / peak, "roofline_tokens_per_s": peak * 1e9 / bytes_tok,
                              "what": "whole-token algorithmic bytes (weights + attended f16 KV at the mean n_kv) x tokens/s vs ONE GPU's "
                                      "measured HBM peak (single stream: stages run one after the other)"}

    if rank == 0 and world == 1 and not args.value_only:
        tok1 = int(ids[-1])
        # ---- prompt batch (llama_decode with n_tokens = 512): its own context, the reference's n_ubatch in one call
        try:
            pc = engine.Context(m, 1024)
            pp = np.random.default_rng(7).integers(0, cfg.n_vocab, size=512).tolist()
            pms = []
            for _ in range(3):
                pc.kv_clear()
                t0 = time.perf_counter()
                pc.decode(pp, 0, want_logits=False)
                torch.cuda.synchronize()
                pms.append(1e3 * (time.perf_counter() - t0))
            pc.close()
            pm = float(np.mean(pms[1:]))
            line["prompt_batch"] = {"tokens": 512, "ms": pm, "tokens_per_s": 512e3 / pm,
                                    "kernels": ("k_quant_batch_mma + k_mma_batch_q80 (block-diagonal fp16 HMMA: exact 4-byte lane sums)" if ftype == "Q8_0" else
                                                "k_quant_batch_mma + k_mma_batch (exact fp16 HMMA per AVX2 lane-slice)") +
                                               " + k_attn_scores_batch / k_attn_softmax_rows / k_attn_pv_batch",
                                    "what": "b200_decode of one 512-token batch at positions 0..511, wall clock incl. the final synchronisation"}
            try:   # tensor-pipe utilisation of the batch mat-mul kernel from the committed ncu --set full capture
                with open(os.path.join(ROOT, "profiles", "prefill_tensor.json")) as f:
                    line["prompt_batch"]["tensor_pipe"] = json.load(f)["k_mma_batch_q80" if ftype == "Q8_0" else "k_mma_batch"]
            except Exception:
                pass
        except Exception as e:
            line["prompt_batch"] = {"error": str(e)}
        # ---- end to end, additive token-level seam: host token in, host logits out, host arg-max
        n_e2e = min(args.steps, 4) * min(burst, 64)
        lg = c.decode([tok1], pos0)
        t0 = time.perf_counter()
        for i in range(n_e2e):
            t = int(np.argmax(lg))
            lg = c.decode([t], pos0 + (i % min(burst, 64)))
        e2e_dt = time.perf_counter() - t0
        decode_e2e = {"value": n_e2e / e2e_dt, "unit": "tokens/s", "h2d_bytes_per_token": 32, "d2h_bytes_per_token": 4 * cfg.n_vocab,
                      "what": "b200_decode(token from host) -> logits to host -> host arg-max, per token"}
        # ---- end to end through the nine bridge symbols (second copy of the model in HBM: the pod's own)
        try:
            n_gen = min(burst, 64)
            n_prompt = (prefill if prefill else ctx - 4 - n_gen)
            if prefill:
                n_gen = 256
            be = bridge_e2e(path, ctx, n_prompt, n_gen, reps=2 if args.config == DEFAULT_CONFIG else 1)
            # per generated token the bridge copies the 32-byte decode state to the device and — the Janus sampler runs on the host,
            # like the reference's — the whole logits row (4 * n_vocab bytes) back; the prompt goes in once as token ids
            line["e2e"] = {"value": be["value"], "unit": "tokens/s", "h2d_bytes_per_step": 32 * be["generated"] + 4 * n_prompt,
                           "d2h_bytes_per_step": 4 * cfg.n_vocab * be["generated"],
                           "what": f"init -> initContext -> doInference(host text prompt of {n_prompt} token ids, predict {n_gen}) -> status poller: "
                                   f"generated tokens / time from the first generated piece to doInference's return (n_kv {n_prompt}..{n_prompt + n_gen})",
                           "doInference": be, "b200_decode": decode_e2e}
        except Exception as e:
            line["e2e"] = dict(decode_e2e, h2d_bytes_per_step=32 * burst, d2h_bytes_per_step=4 * cfg.n_vocab * burst,
                               bridge_error=str(e))
        # ---- live per-kernel roofline (event pair around every launch of an un-graphed token at n_kv ~ ctx)
        acc = {}
        reps = 4
        ppos = pos0 + min(burst, 64) // 2
        for i in range(reps + 1):
            prof = c.profile_token(tok1, ppos)
            if i == 0:
                continue   # warm-up
            for k, (ms, n) in prof.items():
                a = acc.setdefault(k, [0.0, 0]); a[0] += ms; a[1] += n
        types = G.tensor_types(cfg, ftype)
        E, FF = cfg.n_embd, cfg.n_ff
        gu_bytes = np.mean([G.row_bytes(types[f"blk.{i}.ffn_gate.weight"], E) * FF + G.row_bytes(types[f"blk.{i}.ffn_up.weight"], E) * FF
                            for i in range(cfg.n_layer)]) + 4 * E
        gu_ms = acc["gate_up"][0] / acc["gate_up"][1]
        achieved = gu_bytes / (gu_ms * 1e-3) / 1e9
        tot_ms = sum(v[0] for v in acc.values()) / reps
        # DRAM traffic of one gate|up launch from the committed ncu --set full capture (profiles/traffic.json, written by
        # scripts/ncu_summary.py): dram__bytes_read.sum + dram__bytes_write.sum (8B Q4_K_M shapes)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)["gate_up"]
            if args.config == DEFAULT_CONFIG:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "kernel": "k_matvec<EPI_SILU> (ffn gate|up, fused RMSNorm + activation quant + SiLU*mul)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "peak_source": peak_src, "bytes_per_launch": float(gu_bytes), "ms_per_launch": gu_ms,
                            "share_of_token_time": acc["gate_up"][0] / reps / tot_ms}
        # the same kernel outside the token's dependency chain: the gate|up launches of all layers back to back
        # (distinct tiles per launch), one event pair around the lot
        iso_ms, iso_n = c.profile_kind("gate_up", ppos, reps=8)
        line["roofline"]["isolated"] = {"ms_per_launch": iso_ms, "launches_timed": iso_n, "achieved": gu_bytes / (iso_ms * 1e-3) / 1e9,
                                        "frac": gu_bytes / (iso_ms * 1e-3) / 1e9 / peak,
                                        "what": "gate|up launches of every layer back to back (distinct weights per launch, PDL), one CUDA-event pair"}
        line["kernel_ms_per_token"] = {k: round(v[0] / reps, 4) for k, v in acc.items()}
        # ---- the reference's CPU llama_decode on this host, bounded sample (always BASELINE.json configs[0]: 8B Q4_K_M)
        try:
            if args.no_cpu:
                raise RuntimeError("skipped (--no-cpu)")
            cpu_tps, info = cpu_reference_run(model_path(), steps=4, warmup=1)
            line["cpu_baseline"] = dict(info, value=cpu_tps, unit="tokens/s")
        except Exception as e:  # the oracle library failing to load must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    elif world > 1:
        # multi-GPU lines carry an e2e too: the pipeline call is itself host-facing (first token id from the host, the
        # burst's ids back to every rank's host memory) — barrier-bracketed wall clock around the timed calls
        line["e2e"] = {"value": (tokens + args.steps * prefill) / wall, "unit": "tokens/s", "h2d_bytes_per_step": 16 * world,
                       "d2h_bytes_per_step": 4 * burst * world,
                       "what": "b200_pipeline_generate_greedy per burst (host token id in, token ids to host on every rank), wall clock "
                               "between barriers over the timed steps"}
    if rank == 0:
        print(json.dumps(line))
    c.close(); m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
