"""Secondary measurements of SURVEY.md 8(d) on one B200 (the headline line is bench.py's): every BASELINE.json config
through the public API, CUDA events, inputs resident in HBM, rotating input sets larger than L2 where the size allows.

    python scripts/sweep.py [--out profiles/rXX_sweep.json] [--quick]

  cfg1  [4,128,750]   n_q=8   encode + decode
  cfg2  [64,128,750]  n_q=32  encode, decode, eval forward
  cfg3  cfg2 shapes, train(): forward (search + quantized + losses + EMA statistics/apply + expiry) and forward+backward
  cfg4  [32,128,4500] n_q=16  one call, and as the 31 one-second segment calls model.py:141-145 makes
  cfg5  N in {1e4 .. 1e7 (1e8 with --full)} frames x n_q in {2, 8, 32}: frames/s and fraction of the tensor peak
Tensor fraction = frames/s * n_q * 2*K*D / measured bf16 peak (MEASURED_PEAKS.json); decode GB/s counts HBM bytes
(codes in + fp32 out) and, separately, the L2 table gathers.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import encodec_pytorch_b200 as E  # noqa: E402

D, K = 128, 1024
FLOP = 2 * K * D


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["bf16_tflops"]), float(j["hbm_gbs"])
    return 1590.0, 6650.0


def latents(b, t, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, D, t, generator=g, dtype=torch.float32).to(dev)


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def quantizer(n_q, dev, kmeans=False):
    torch.manual_seed(0)
    return E.ResidualVectorQuantizer(dimension=D, n_q=n_q, bins=K, kmeans_init=kmeans).to(dev)


def encode_decode(name, b, t, n_q, fr, bw, dev, reps, out):
    tf_peak, hbm_peak = peaks()
    q = quantizer(n_q, dev).eval()
    frames = b * t
    nsets = max(1, min(8, int(2.0e8 // (frames * (4 * D + 8 * n_q))) + 1))
    xs = [latents(b, t, 1234 + i, dev) for i in range(nsets)]
    with torch.no_grad():
        ms_enc = timed(lambda i: q.encode(xs[i % nsets], fr, bw), reps)
        codes = q.encode(xs[0], fr, bw)
        ms_dec = timed(lambda i: q.decode(codes), reps)
        ms_fwd = timed(lambda i: q(xs[i % nsets], fr, bw), reps)
    nq_eff = int(codes.shape[0])
    out[name] = {
        "shape": [b, D, t], "n_q": nq_eff, "frames": frames,
        "encode_ms": ms_enc, "encode_frames_per_s": frames / ms_enc * 1e3,
        "encode_tensor_frac": frames / ms_enc * 1e3 * nq_eff * FLOP / (tf_peak * 1e12),
        "decode_ms": ms_dec, "decode_frames_per_s": frames / ms_dec * 1e3,
        "decode_hbm_gbs": frames * (8 * nq_eff + 4 * D) / ms_dec / 1e6,
        "decode_l2_gather_gbs": frames * nq_eff * 4 * D / ms_dec / 1e6,
        "eval_forward_ms": ms_fwd, "eval_forward_frames_per_s": frames / ms_fwd * 1e3,
        "input_sets": nsets,
    }
    print(name, json.dumps(out[name]), flush=True)


def training(name, b, t, n_q, fr, bw, dev, reps, out):
    """cfg3: training forward with EMA update + dead-code expiry (random-init codebooks), then forward+backward."""
    _, hbm_peak = peaks()
    q = quantizer(n_q, dev).train()
    frames = b * t
    xs = [latents(b, t, 77 + i, dev) for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            ms_fwd = timed(lambda i: q(xs[i % 4], fr, bw), reps)

        def fb(i):
            x = xs[i % 4].detach().requires_grad_(True)
            r = q(x, fr, bw)
            (r.quantized.sum() + r.penalty).backward()
        ms_fb = timed(fb, reps)
        # k-means init step (first training forward of a kmeans_init=True stack): 50 Lloyd iterations per stage
        # (twice: the first k-means of a shape in a process also captures the Lloyd iteration's CUDA graph)
        ms_km_all = []
        for _ in range(2):
            qk = quantizer(min(n_q, 8), dev, kmeans=True).train()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            with torch.no_grad():
                qk(xs[0], fr, bw * min(n_q, 8) / n_q)
            e.record()
            torch.cuda.synchronize()
            ms_km_all.append(a.elapsed_time(e))
        ms_km_first, ms_km = ms_km_all
    ema_bytes = frames * n_q * (4 * D + 8) + n_q * (3 * K * D + 2 * K) * 4
    out[name] = {
        "shape": [b, D, t], "n_q": n_q, "frames": frames,
        "train_forward_ms": ms_fwd, "train_forward_frames_per_s": frames / ms_fwd * 1e3,
        "train_forward_backward_ms": ms_fb, "train_forward_backward_frames_per_s": frames / ms_fb * 1e3,
        "ema_algorithmic_bytes_per_step": ema_bytes,
        "kmeans_init_step_ms": ms_km, "kmeans_init_first_call_ms": ms_km_first, "kmeans_init_stages": min(n_q, 8), "kmeans_iters": 50,
    }
    print(name, json.dumps(out[name]), flush=True)


def segmented(name, b, t_total, seg, n_q, fr, bw, dev, reps, out):
    """cfg4 the way model.py:141-145 drives the 48 kHz model: one RVQ call per 1-s segment (31 calls, last one shorter)."""
    q = quantizer(n_q, dev).eval()
    x = latents(b, t_total, 5, dev)
    segs = [x[:, :, o:o + seg] for o in range(0, t_total, seg)]
    # the segments as separate tensors (each one comes out of its own SEANet call in model.py:141-145)
    segs = [s.contiguous() for s in segs]
    with torch.no_grad():
        ms = timed(lambda i: [q.encode(s, fr, bw) for s in segs], reps)
        ms_batched = timed(lambda i: q.encode_segments(segs, fr, bw), reps)
        same = all(torch.equal(a, c) for a, c in zip(q.encode_segments(segs, fr, bw), [q.encode(s, fr, bw) for s in segs]))
    out[name] = {"shape": [b, D, t_total], "segment_frames": seg, "calls": len(segs), "n_q": n_q,
                 "encode_ms_all_segments": ms, "encode_frames_per_s": b * t_total / ms * 1e3,
                 "encode_segments_ms": ms_batched, "encode_segments_frames_per_s": b * t_total / ms_batched * 1e3,
                 "encode_segments_equal_to_loop": bool(same)}
    print(name, json.dumps(out[name]), flush=True)


def trained_like(name, b, t, n_q, fr, bw, dev, reps, out, steps=25):
    """codebooks fitted to the latents (k-means init on the first batch, then `steps` EMA training steps on
    fresh batches) so that residual norms decay stage by stage as in a trained model; then the eval encode is timed."""
    from encodec_pytorch_b200 import _ops as ops
    tf_peak, _ = peaks()
    q = quantizer(n_q, dev, kmeans=True).train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            for i in range(steps):
                q(latents(b, t, 500 + i, dev), fr, bw)
    q.eval()
    xs = [latents(b, t, 900 + i, dev) for i in range(6)]
    with torch.no_grad():
        ms = timed(lambda i: q.encode(xs[i % 6], fr, bw), reps)
        with ops.search_counters(dev) as counters:
            codes = q.encode(xs[0], fr, bw)
        st = counters.read()
        r = xs[0]
        norms = []
        for i in (0, 1, 3, 7, 15, n_q - 1):
            part = q.decode(codes[: i + 1])
            norms.append(float((xs[0] - part).norm() / xs[0].norm()))
    frames = b * t
    out[name] = {"shape": [b, D, t], "n_q": n_q, "training_steps": steps, "encode_ms": ms,
                 "encode_frames_per_s": frames / ms * 1e3, "encode_tensor_frac": frames / ms * 1e3 * n_q * FLOP / (tf_peak * 1e12),
                 "certified_share": st["certified"] / max(1, st["searched"]), "rescored": st["rescored"], "fullscan": st["fullscan"],
                 "relative_residual_norm_after_stage_1_2_4_8_16_last": norms}
    print(name, json.dumps(out[name]), flush=True)


def bitpack(name, b, t, n_q, dev, reps, out):
    """8(f)-1: the byte streams of binary.BitPacker for the codes of one encode (10 bits per code), and back."""
    from encodec_pytorch_b200 import binary as BN
    _, hbm_peak = peaks()
    q = quantizer(n_q, dev).eval()
    with torch.no_grad():
        codes = q.encode(latents(b, t, 11, dev), 75, None)
    frame = codes.transpose(0, 1)                        # [B, K, T] view of the search's [K, B, T] output (model.py:166)
    ms_p = timed(lambda i: BN.pack_frame(frame, 10), reps)
    packed = BN.pack_frame(frame, 10)
    ms_u = timed(lambda i: BN.unpack_frame(packed, n_q, t, 10), reps)
    vals = b * t * n_q
    algo = vals * 8 + packed.numel()                     # int64 codes + packed bytes (either direction)
    out[name] = {"shape_bkt": [b, n_q, t], "bits": 10, "values": vals, "packed_bytes": int(packed.numel()),
                 "pack_ms": ms_p, "pack_hbm_gbs": algo / ms_p / 1e6, "unpack_ms": ms_u, "unpack_hbm_gbs": algo / ms_u / 1e6,
                 "hbm_peak_gbs": hbm_peak}
    print(name, json.dumps(out[name]), flush=True)


def bulk(dev, out, full):
    """cfg5: frames/s and tensor fraction vs frame count."""
    tf_peak, _ = peaks()
    rows = []
    sizes = [10_000, 100_000, 1_000_000, 10_000_000] + ([100_000_000] if full else [])
    for n_q in (2, 8, 32):
        q = quantizer(n_q, dev).eval()
        for n in sizes:
            b = max(1, round(n / 750))
            chunk_b = min(b, 13334)                     # <= 1e7 frames per call: 5.1 GB of latents + codes per call
            x = latents(min(chunk_b, 2000), 750, 3, dev)
            if chunk_b > x.shape[0]:
                x = x.repeat((chunk_b + x.shape[0] - 1) // x.shape[0], 1, 1)[:chunk_b].contiguous()
            calls = (b + chunk_b - 1) // chunk_b
            reps = max(2, min(50, int(3e7 // (chunk_b * 750 * calls))))
            with torch.no_grad():
                ms = timed(lambda i: [q.encode(x, 75, None) for _ in range(calls)], reps, warm=2)
            frames = chunk_b * 750 * calls
            fps = frames / ms * 1e3
            rows.append({"frames": frames, "n_q": n_q, "calls": calls, "ms": ms, "frames_per_s": fps,
                         "tensor_frac": fps * n_q * FLOP / (tf_peak * 1e12)})
            print("cfg5", json.dumps(rows[-1]), flush=True)
            del x
            torch.cuda.empty_cache()
    out["cfg5_bulk"] = rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--full", action="store_true", help="include the 1e8-frame leg of cfg5")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a B200"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    reps = 5 if args.quick else 30
    tf_peak, hbm_peak = peaks()
    out = {"peaks": {"bf16_tflops": tf_peak, "hbm_gbs": hbm_peak}, "device": torch.cuda.get_device_name(0)}
    encode_decode("cfg1_24khz_6kbps", 4, 750, 8, 75, 6.0, dev, reps, out)
    encode_decode("cfg2_24khz_24kbps", 64, 750, 32, 75, 24.0, dev, reps, out)
    training("cfg3_training", 64, 750, 32, 75, 24.0, dev, max(3, reps // 3), out)
    encode_decode("cfg4_48khz_one_call", 32, 4500, 16, 150, 24.0, dev, reps, out)
    segmented("cfg4_48khz_31_segments", 32, 4500, 150, 16, 150, 24.0, dev, max(3, reps // 3), out)
    trained_like("cfg2_trained_like", 64, 750, 32, 75, 24.0, dev, reps, out)
    bitpack("bitpack_cfg2_codes", 64, 750, 32, dev, reps, out)
    bitpack("bitpack_cfg4_codes", 32, 4500, 16, dev, reps, out)
    bitpack("bitpack_16M_codes", 512, 1000, 32, dev, reps, out)
    if not args.quick:
        bulk(dev, out, args.full)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
