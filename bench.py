#!/usr/bin/env python
"""Benchmark of the IMS-Toucan hot path on B200: audio-seconds/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on): BigVGAN generator
alone, batch 64 synthetic 80-bin mel spectrograms x 500 frames -> 24 kHz wave (512 s of audio per step per
GPU).  `--workload acoustic` is configs[2] (ToucanTTS, 128 ragged phoneme sequences <= 200 tokens -> mel) and
`--workload e2e` the per-GPU shard of configs[3] (64 utterances text -> wave); both are extra measurements,
not the headline.  One process per GPU; under torchrun every rank synthesises its own shard (weak scaling, no
collective on the data path; NCCL only gathers output lengths and the max-over-ranks time).

Prints ONE JSON line (see the contract in the task description): `value` is device-timed with the
mels resident in HBM, `e2e` goes through the module's public batched call with pinned HOST buffers
(H2D of the mels and D2H of the waveforms inside the timed region), `roofline` is the tensor-core
roofline of the conv kernel, `cpu_baseline` is the oracle port of the reference timed on this box's
host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_FRAME = 648_241_152          # 2*MAC of every Conv1d/ConvTranspose1d of the generator
SAMPLES_PER_FRAME = 384
SAMPLE_RATE = 24_000
METRIC = "audio-seconds/sec"
UNIT = "audio_s/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
            return
        # wait (bounded) for the first sample: nvidia-smi has then finished its start-up
        t_end = time.time() + 5.0
        while time.time() < t_end:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                break
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons))
        return out


def build_generator(kind, precision, device, activation_dtype="f32"):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory  # weights factory only (state_dict values); not on the timed path
    sd = factory.make_state_dict(kind, 1234)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "g.pt")
        torch.save({"generator": sd}, path)
        cls = tb.BigVGAN if kind == "bigvgan" else tb.HiFiGANGenerator
        model = cls(path, precision=precision, activation_dtype=activation_dtype).to(device)
    model.remove_weight_norm()
    return model, sd


def cpu_reference_step(kind, fsd, mel):
    from oracle import restate
    fwd = restate.bigvgan_forward if kind == "bigvgan" else restate.hifigan_forward
    with torch.inference_mode():
        return fwd(fsd, mel)


def time_cpu(kind, sd, frames, repeats, warmup):
    """Oracle port of the reference on the host cores; sequential batch-1 like read_to_file
    (ToucanTTSInterface.py:269-280).  Returns audio-s/s and the sample description."""
    from oracle import factory, restate
    torch.set_num_threads(os.cpu_count() or 1)
    fsd = restate.fold_weight_norm(sd)
    mel = factory.make_mel(1, frames, seed=2)[0]
    for _ in range(warmup):
        cpu_reference_step(kind, fsd, mel)
    t0 = time.perf_counter()
    for _ in range(repeats):
        cpu_reference_step(kind, fsd, mel)
    mean = (time.perf_counter() - t0) / repeats     # mean of the timed steps, like the engine arm
    audio = frames * SAMPLES_PER_FRAME / SAMPLE_RATE
    return audio / mean, mean, f"1 utterance x {frames} frames ({audio:.1f} s audio), batch-1, mean of {repeats}"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    Python and does not travel to the GPU box) with all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import factory
    sd = factory.make_state_dict(args.vocoder, 1234)
    cores = os.cpu_count() or 1
    val, best, sample = time_cpu(args.vocoder, sd, args.frames, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(best * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"{'BigVGAN' if args.vocoder == 'bigvgan' else 'HiFiGAN'} generator alone: batch {args.batch} "
                        f"synthetic 80-bin mels x {args.frames} frames -> 24 kHz wave (per GPU)",
            "batch_per_gpu": args.batch, "frames": args.frames, "vocoder": args.vocoder,
            "operand_precision": args.precision, "residual_stream": args.activations, "weights": "random-init (oracle.factory seed 1234)",
            "l2": "working set per step (>6 GB of activations) exceeds the 126 MB L2; no explicit flush"}


def tts_flops(t_list, f_list, with_vocoder):
    """Algorithmic FLOPs of the acoustic model: 24.5 M per phoneme + 63.8 M per frame +
    9 216 (T^2 + F^2) attention; plus 648.24 M per frame for the vocoder."""
    fl = 0.0
    for t, f in zip(t_list, f_list):
        fl += 24.5e6 * t + 63.8e6 * f + 9216.0 * (t * t + f * f)
        if with_vocoder:
            fl += FLOP_PER_FRAME * (2 * (f // 2))
    return fl


def run_tts_workload(args, rank, local_rank, world, dev, dist):
    """configs[2] (acoustic: 128 ragged utterances -> mel) and the per-GPU shard of configs[3] (e2e: 64 utterances
    text -> wave through BigVGAN/HiFiGAN).  Same JSON contract as the vocoder workload."""
    import random

    import ims_toucan_prosody_variance_b200 as tb
    from ims_toucan_prosody_variance_b200 import ops
    from oracle import factory, restate
    e2e_mode = args.workload == "e2e"
    n_utt = args.batch if args.batch_given else (64 if e2e_mode else 128)
    tsd = factory.make_state_dict("toucantts", 1234)
    tts = tb.ToucanTTS(weights=tsd, precision=args.acoustic_precision).to(dev)
    tts.store_inverse_all()
    voc = vsd = None
    if e2e_mode:
        voc, vsd = build_generator(args.vocoder, args.precision, dev)
    rng = random.Random(3 + rank)
    lens = [rng.randint(20, 200) for _ in range(n_utt)]
    t_max = max(lens)
    text_host = torch.zeros((n_utt, t_max, 62), dtype=torch.float32)
    for i, n in enumerate(lens):
        text_host[i, :n] = factory.make_phoneme_tensor(n, 1000 * rank + i)
    text_host = text_host.pin_memory()
    emb = torch.stack([factory.make_utterance_embedding(i) for i in range(n_utt)]).to(dev)
    tlen = torch.tensor(lens, dtype=torch.int32)
    lang = torch.full((n_utt,), 12, dtype=torch.int64)
    text_dev = text_host.to(dev)
    eng = tb.TextToWave(tts, voc) if e2e_mode else None

    def step(text):
        if e2e_mode:
            wave, wlen, r = eng.synthesize_padded(text, tlen, emb, lang_ids=lang, noise="device")
            return wave, wlen, r
        r = tts.synthesize_batch(text, tlen, utterance_embedding=emb, lang_ids=lang, noise="device")
        return r["mel_ncl"], r["mel_lengths"], r

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the clock sampler (an nvidia-smi process) starts BEFORE the warm-up: its start-up takes driver locks for tens of
    # milliseconds, which must not land inside the timed region; samples are taken through warm-up and timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out, out_len, r = step(text_dev)
    barrier()
    frames = [int(v) for v in r["frames_host"]]
    audio_s = sum(2 * (f // 2) for f in frames) * SAMPLES_PER_FRAME / SAMPLE_RATE
    launches0 = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out, out_len, r = step(text_dev)
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    launches = (ops.LAUNCHES - launches0) // args.steps
    out_host = torch.empty(out.shape, dtype=torch.float32).pin_memory()   # durations are noise-independent: fixed shape

    def e2e_step():
        t = text_host.to(dev, non_blocking=True)
        o, _, _ = step(t)
        out_host.copy_(o, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    # the only collective: output lengths + whole-job audio seconds
    from ims_toucan_prosody_variance_b200 import sharding
    all_len = sharding.gather_output_lengths(range(rank * n_utt, (rank + 1) * n_utt), out_len.to(torch.int64), world * n_utt, device=dev)
    total_audio, _ = sharding.reduce_metrics(audio_s, ms_step, device=dev)
    assert int(all_len.sum()) > 0 and bool(torch.isfinite(out[0, ..., :int(out_len[0])]).all())
    if rank != 0:
        return
    value = total_audio / (ms_step / 1e3)
    tflops_peak, hbm_peak, peak_src = peaks()
    flops = tts_flops(lens, frames, e2e_mode)
    achieved = flops / (ms_step / 1e3) / 1e12
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        tf = restate.fold_weight_norm(tsd)
        vf = restate.fold_weight_norm(vsd) if e2e_mode else None
        text1 = factory.make_phoneme_tensor(100, 1)
        best, fr = float("inf"), 0
        for _ in range(2):
            t0 = time.perf_counter()
            with torch.inference_mode():
                ref = restate.toucantts_forward(tf, text1, factory.make_utterance_embedding(1), lang_id=12)
                if e2e_mode:
                    (restate.bigvgan_forward if args.vocoder == "bigvgan" else restate.hifigan_forward)(vf, ref["mel"].t())
            best = min(best, time.perf_counter() - t0)
            fr = ref["mel"].shape[0]
        cpu = {"value": round(fr * SAMPLES_PER_FRAME / SAMPLE_RATE / best, 3), "unit": UNIT, "cores": os.cpu_count() or 1,
               "kind": "port", "sample": f"1 utterance x 100 phonemes ({fr} frames), batch-1, best of 2"}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.acoustic_precision, "data": "synthetic",
        "config": {"workload": (f"text->wave, {n_utt} ragged utterances (20..200 phonemes) per GPU through ToucanTTS + {args.vocoder}"
                                if e2e_mode else f"ToucanTTS acoustic model only: {n_utt} ragged phoneme sequences (20..200 tokens) -> mel, per GPU"),
                   "utterances_per_gpu": n_utt, "phonemes": sum(lens), "frames": sum(frames), "acoustic_precision": args.acoustic_precision,
                   "vocoder": args.vocoder if e2e_mode else None, "weights": "random-init (oracle.factory seed 1234, calibrated duration head)",
                   "l2": "activations per step exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": round(total_audio / (ms_e2e / 1e3), 2), "unit": UNIT, "h2d_bytes_per_step": text_host.numel() * 4,
                "d2h_bytes_per_step": out.numel() * 4, "ms_per_step": round(ms_e2e, 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": tflops_peak, "unit": "TFLOP/s",
                     "frac": round(achieved / tflops_peak, 4), "traffic": None, "peak_source": peak_src,
                     "kernel": "whole step (conv1d_umma + acoustic kernels); algorithmic FLOPs per SURVEY.md 8d"},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def run_config4(args, rank, world, dev, dist):
    """BASELINE.json configs[3]: ONE global batch of 512 ragged utterances (20..200 phonemes), partitioned by utterance
    over the ranks with sharding.partition_lpt (longest-processing-time first on the estimated cost), each rank running
    text -> wave through TextToWave (length-sorted batches of <= 128 utterances) with no collective on the data path; the ranks
    then exchange output lengths and timings.  Strong scaling: the total work is fixed as N grows.  Returns the extra
    `config4` object of the bench line (on every rank; rank 0 prints it)."""
    import random

    import ims_toucan_prosody_variance_b200 as tb
    from ims_toucan_prosody_variance_b200 import sharding
    from oracle import factory
    n_total = args.config4_utterances
    rng = random.Random(4)
    lens = [rng.randint(20, 200) for _ in range(n_total)]
    shards = sharding.partition_lpt([sharding.estimate_cost(n) for n in lens], world)
    mine = shards[rank]
    tts = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234), precision=args.acoustic_precision).to(dev)
    tts.store_inverse_all()
    if not args.config4_no_graphs:
        tts.enable_cuda_graphs()       # both acoustic segments replayed per padded batch shape (captured by the warm-up pass)
    voc, _ = build_generator(args.vocoder, args.precision, dev, args.activations)
    kw = {}
    if args.config4_max_batch:
        kw["max_batch"] = args.config4_max_batch
    if args.config4_padding_ratio:
        kw["max_padding_ratio"] = args.config4_padding_ratio
    eng = tb.TextToWave(tts, voc, **kw)
    texts = [factory.make_phoneme_tensor(lens[i], 5000 + i) for i in mine]
    emb = torch.stack([factory.make_utterance_embedding(i) for i in mine]) if mine else torch.zeros((0, 64))
    lang = torch.full((len(mine),), 12, dtype=torch.int64)

    def one_pass():
        return eng.synthesize(texts, emb, lang_ids=lang, noise="device", device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    waves = one_pass()                      # warm-up (workspaces, planner caches)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    ev0.record()
    for _ in range(reps):
        waves = one_pass()
    ev1.record()
    barrier()
    my_ms = ev0.elapsed_time(ev1) / reps
    wlen = [int(w.numel()) for w in waves]
    audio = sum(wlen) / SAMPLE_RATE
    all_len = sharding.gather_output_lengths(mine, wlen, n_total, device=dev)
    times = torch.zeros(world, dtype=torch.float64, device=dev)
    times[rank] = my_ms
    audios = torch.zeros(world, dtype=torch.float64, device=dev)
    audios[rank] = audio
    if dist is not None:
        dist.all_reduce(times)
        dist.all_reduce(audios)
    t_max, t_mean = float(times.max()), float(times.mean())
    total_audio = float(audios.sum())
    assert int((all_len > 0).sum()) == n_total and abs(float(all_len.sum()) / SAMPLE_RATE - total_audio) < 1e-6 * total_audio + 1e-3
    return {"workload": f"text->wave, ONE global batch of {n_total} ragged utterances (20..200 phonemes) partitioned by utterance "
                        f"(LPT on estimated cost) over {world} GPU(s), ToucanTTS ({args.acoustic_precision}"
                        f"{'' if args.config4_no_graphs else ', CUDA-graph replay per batch shape'}) + {args.vocoder}",
            "scaling": "strong", "utterances": n_total, "audio_s": round(total_audio, 2), "ms": round(t_max, 3),
            "value": round(total_audio / (t_max / 1e3), 2), "unit": UNIT, "rank_ms_max": round(t_max, 3),
            "rank_ms_mean": round(t_mean, 3), "imbalance": round(t_max / t_mean - 1.0, 4),
            "rank_audio_s": [round(float(v), 1) for v in audios.tolist()]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--vocoder", default="bigvgan", choices=["bigvgan", "hifigan"])
    ap.add_argument("--precision", default="f16", choices=["f16", "tf32", "fp32"])
    ap.add_argument("--workload", default="vocoder", choices=["vocoder", "acoustic", "e2e"])
    # fp16 operands with fp32 accumulation: the mantissa of tf32 at twice the MMA rate and half the operand bytes; held
    # to the same parity bound as tf32 (mel rel-L1 <= 1e-3; measured 1.9e-4 for both, tests/test_toucantts_gpu.py)
    ap.add_argument("--acoustic-precision", default="f16", choices=["tf32", "f16", "fp32"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--activations", default="f16", choices=["f32", "f16"],
                    help="storage type of the vocoder residual stream in HBM (f16: fused residual pairs; f32: every conv its own launch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the extra sharded 512-utterance text->wave run")
    ap.add_argument("--config4-utterances", type=int, default=512)
    ap.add_argument("--config4-max-batch", type=int, default=None, help="TextToWave bucket size (default: the engine's)")
    ap.add_argument("--config4-padding-ratio", type=float, default=None, help="TextToWave longest/shortest ratio per bucket")
    ap.add_argument("--config4-only", action="store_true", help="debug: print only the config4 object")
    ap.add_argument("--config4-no-graphs", action="store_true", help="acoustic model launched eagerly instead of CUDA-graph replay")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 64
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.config4_only:
        c4 = run_config4(args, rank, world, dev, dist)
        if rank == 0:
            print(json.dumps(c4), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return
    if args.workload != "vocoder":
        run_tts_workload(args, rank, local_rank, world, dev, dist)
        if dist is not None:
            dist.destroy_process_group()
        return

    from ims_toucan_prosody_variance_b200 import ops
    from oracle import factory
    model, sd = build_generator(args.vocoder, args.precision, dev, args.activations)
    mel_host = factory.make_mel(args.batch, args.frames, seed=100 + rank).pin_memory()
    mel_dev = mel_host.to(dev)
    lengths = torch.full((args.batch,), args.frames, dtype=torch.int32)
    n_samples = args.frames * SAMPLES_PER_FRAME
    wave_host = torch.empty((args.batch, n_samples), dtype=torch.float32).pin_memory()
    audio_s_per_step = args.batch * n_samples / SAMPLE_RATE

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ----
    # the clock sampler (an nvidia-smi process) starts BEFORE the warm-up: its start-up takes driver locks for tens of
    # milliseconds, which must not land inside the timed region; samples are taken through warm-up and timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        model.forward_batch(mel_dev, lengths)
    barrier()
    launches0 = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        wave = model.forward_batch(mel_dev, lengths)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    launches = (ops.LAUNCHES - launches0) // args.steps
    ms_step = ms_total / args.steps
    value = world * audio_s_per_step / (ms_step / 1e3)

    # ---- end to end through the public call with host buffers ----
    def e2e_step():
        m = mel_host.to(dev, non_blocking=True)
        w = model.forward_batch(m, lengths)
        wave_host.copy_(w, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    e2e_value = world * audio_s_per_step / (ms_e2e / 1e3)

    # ---- NCCL: gather output lengths (the only collective of the path) ----
    out_lengths = (lengths.to(dev) * SAMPLES_PER_FRAME).to(torch.int64)
    if dist is not None:
        gathered = [torch.empty_like(out_lengths) for _ in range(world)]
        dist.all_gather(gathered, out_lengths)
        total_samples = int(sum(int(g.sum()) for g in gathered))
    else:
        total_samples = int(out_lengths.sum())
    assert total_samples == world * args.batch * n_samples
    finite = bool(torch.isfinite(wave).all()) and float(wave.abs().max()) <= 1.0
    if not finite:
        raise SystemExit("bench.py: generator produced non-finite or out-of-range samples")
    # ---- extra (after the headline numbers are taken): configs[3], one global batch sharded by utterance ----
    config4 = None
    if not args.no_config4:
        del wave
        model._buffers_cache = {}
        torch.cuda.empty_cache()
        config4 = run_config4(args, rank, world, dev, dist)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    tflops_peak, hbm_peak, peak_src = peaks()
    flops_step = FLOP_PER_FRAME * args.batch * args.frames
    achieved = flops_step / (ms_step / 1e3) / 1e12  # per GPU: each rank does the same work in ms_step
    traffic = None   # DRAM bytes per launch of the same kernel, from the committed ncu pass (profiles/)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.vocoder}_b{args.batch}_f{args.frames}_{args.precision}")
    roofline = {"bound": "tensor", "achieved": round(achieved, 2), "peak": tflops_peak, "unit": "TFLOP/s",
                "frac": round(achieved / tflops_peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": "conv1d_umma_kernel + respair_kernel (the two tcgen05 implicit-GEMM conv kernels; a fused residual pair is "
                          "one respair launch): the step is their launches back to back, so achieved = (algorithmic FLOPs per "
                          f"launch = {FLOP_PER_FRAME} per mel frame x {args.batch * args.frames} frames / {int(launches)} launches) / "
                          "(average launch duration = CUDA-event step time / launches)",
                "launch_ms_avg": round(ms_step / max(int(launches), 1), 4)}

    cpu = None
    if not args.no_cpu_baseline:
        v, best, sample = time_cpu(args.vocoder, sd, args.frames, 2, 1)
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"f16": "fp16", "tf32": "tf32", "fp32": "fp32"}[args.precision], "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": mel_host.numel() * 4,
                "d2h_bytes_per_step": wave_host.numel() * 4, "ms_per_step": round(ms_e2e, 4)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "vocoder_rtf": round((ms_step / 1e3) / audio_s_per_step, 8),
        "config4": config4,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
