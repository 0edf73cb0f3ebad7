#!/usr/bin/env python
"""Benchmark of the MEM pretraining hot path on B200 (contract in the task prompt).

    python bench.py --gpus 1 --steps K --warmup W             # our arm (CUDA, libmemb)
    python bench.py --impl reference --gpus 1 --steps K ...   # reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                        # one rank per GPU

Prints ONE JSON line on rank 0.  Workloads:

  histogram  event -> polarity histogram rasterisation, 10M uniform events at the
             N-ImageNet 640x480 sensor (largest single-GPU case of BASELINE config 2)
  finetune   ViT-B/16 ft_vit classification step, batch 128/GPU (BASELINE config 5)
  event_pipeline  one training batch of raw streams -> model input: fused event
             augmentations + rasteriser + crop / hot-pixel filter / normalise
  pretrain   ViT-B/16 MEM pretraining step, batch 128/GPU (BASELINE config 3)

`--workload auto` picks `pretrain` when the training-step kernels are built in,
else `histogram`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "histogram", "pretrain", "event_pipeline", "raw_histogram", "finetune"])
    ap.add_argument("--events", type=int, default=10_000_000)
    ap.add_argument("--sensor", default="640x480")
    ap.add_argument("--dist", default="uniform", choices=["uniform", "edge8", "edge50", "hot"],
                    help="histogram workload: spatial distribution of the synthetic stream")
    ap.add_argument("--no-histogram", action="store_true",
                    help="pretrain workload: skip the `histogram` sub-record (the second metric of BASELINE.json)")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--model", default="base", choices=["base", "large"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(key):
    """DRAM bytes per launch of a roofline kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]["bytes"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ----------------------------------------------------------------------------- histogram workload
class HistogramWorkload:
    metric = "Gevents/s histogram rasterise"
    unit = "Gevents/s"
    dtype = "u8"

    def __init__(self, args):
        w, h = (int(v) for v in args.sensor.split("x"))
        self.H, self.W, self.n, self.C = h, w, args.events, 3
        self.kind = getattr(args, "dist", "uniform")
        self.config = {"workload": f"event->histogram rasterise, {self.kind} synthetic stream, {self.n} events, "
                                   f"{w}x{h} sensor (N-ImageNet), float64[N,4] rows -> uint8[H,W,3]",
                       "events": self.n, "sensor_wxh": [w, h], "channels": self.C,
                       "l2_policy": "input (32 B/event, %.0f MB) larger than the 126 MB L2" % (self.n * 32 / 1e6),
                       "parallelism": "independent streams per GPU (no collective)"}

    def make_events(self, seed, n=None, kind=None):
        """SURVEY.md 8(d) streams: uniform; edge-like (events on `segments` random line segments, sigma = 1 px:
        "edge8" is the concentrated case of tools/hist_sweep.py, "edge50" the survey's); hot-pixel (1 % of the events
        on 16 fixed pixels).  t sorted ascending, p = +-1."""
        import numpy as np
        n = n or self.n
        kind = kind or self.kind
        rng = np.random.default_rng(seed)
        ev = np.empty((n, 4), dtype=np.float64)
        if kind.startswith("edge"):
            k = int(kind[4:] or 8)
            seg = rng.integers(0, k, n)
            x0, y0 = rng.uniform(0, self.W, k), rng.uniform(0, self.H, k)
            dx, dy = rng.uniform(-1, 1, k), rng.uniform(-1, 1, k)
            s = rng.uniform(0, min(self.H, self.W) / 2, n)
            ev[:, 0] = np.floor(np.clip(x0[seg] + dx[seg] * s + rng.normal(0, 1, n), 0, self.W - 1))
            ev[:, 1] = np.floor(np.clip(y0[seg] + dy[seg] * s + rng.normal(0, 1, n), 0, self.H - 1))
        else:
            ev[:, 0] = rng.integers(0, self.W, n)
            ev[:, 1] = rng.integers(0, self.H, n)
            if kind == "hot":
                hot = rng.random(n) < 0.01
                hx, hy = rng.integers(0, self.W, 16), rng.integers(0, self.H, 16)
                pick = rng.integers(0, 16, n)
                ev[hot, 0], ev[hot, 1] = hx[pick[hot]], hy[pick[hot]]
        ev[:, 2] = np.sort(rng.uniform(0, 3e5, n))
        ev[:, 3] = rng.integers(0, 2, n) * 2.0 - 1.0
        return ev

    # --- ours
    def setup(self, torch, rank):
        from mem_b200 import _lib
        from mem_b200.process_data import histogram
        self.torch, self._lib, self.histogram = torch, _lib, histogram
        ev = self.make_events(rank)
        self.ev_host = torch.from_numpy(ev).pin_memory()
        self.ev_dev = self.ev_host.cuda(non_blocking=True)
        self.out = None
        self.units_per_step = self.n
        self.h2d, self.d2h = self.n * 32, self.H * self.W * self.C

    def step_device(self):
        self.out = self.histogram(self.ev_dev, self.H, self.W, channels=self.C, check=False)

    def step_e2e(self):
        # public API with HOST input (pinned) and the result read back to the host
        dev = self.ev_host.to("cuda", non_blocking=True)
        out = self.histogram(dev, self.H, self.W, channels=self.C, check=True)
        return out.cpu()

    def verify(self):
        import numpy as np
        from oracle.histogram_ref import event_hist_ref
        n = min(self.n, 2_000_000)
        got = self.histogram(self.ev_dev[:n], self.H, self.W, channels=self.C).cpu().numpy()
        return bool(np.array_equal(got, event_hist_ref(self.ev_host[:n].numpy(), self.H, self.W)))

    def roofline(self, ms_per_step):
        hbm, _, _, how = measured_peaks()
        alg_bytes = 32.0 * self.n + self.C * self.H * self.W
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        # what MEMB_HIST_AUTO picks (csrc/hist.cu make_plan)
        if self.H * self.W <= 51200 and self.n >= (1 << 18):
            kernel = "hist_private (+hist_private_finalize: whole step timed)"
        elif self.n >= (1 << 20) and self.H * self.W <= 64 * 16384:
            kernel = ("adaptive chain, whole step timed: hist_hybrid_prepare (zero-fill + sampled concentration estimate) -> "
                      "hist_scatter_global (spread-out streams: L2 REDs) | hist_hybrid (concentrated streams: hot granules in "
                      "shared memory) -> hist_hybrid_finalize")
        else:
            kernel = "hist_scatter_global (+init, finalize: whole step timed)"
        return {"bound": "hbm", "kernel": kernel,
                "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": ncu_traffic("hist_chain_10M_640x480_" + self.kind) if (self.n, self.W, self.H) == (10_000_000, 640, 480) else None}

    # --- reference's CPU path (numpy np.add.at restatement, oracle/histogram_ref.py)
    def cpu_once(self, ev):
        from oracle.histogram_ref import event_hist_ref
        return event_hist_ref(ev, self.H, self.W)

    def cpu_baseline(self):
        n = min(self.n, 4_000_000)
        ev = self.make_events(0, n)
        self.cpu_once(ev[:100_000])
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 4.0:
            self.cpu_once(ev)
            reps += 1
            if time.perf_counter() - t0 > 25:
                break
        dt = (time.perf_counter() - t0) / reps
        return {"value": round(n / dt / 1e9, 6), "unit": self.unit, "cores": 1, "kind": "port",
                "sample": f"{reps} passes of {n} events ({self.W}x{self.H}) through oracle/histogram_ref.py "
                          f"(np.add.at is serial: 1 thread)"}


# ----------------------------------------------------------------------------- raw recording workload
class RawHistogramWorkload(HistogramWorkload):
    """N-Caltech101 raw 5-byte records -> polarity histogram without float64 rows (decode fused into the scatter)."""
    metric = "Gevents/s raw-record rasterise (N-Caltech101 5-byte records)"

    def __init__(self, args):
        self.H, self.W, self.n, self.C = 180, 240, args.events, 3
        self.config = {"workload": f"N-Caltech101 .bin records (5 B/event) -> uint8[180,240,3], uniform synthetic recording, "
                                   f"{self.n} events; decode fused into the rasteriser",
                       "events": self.n, "sensor_wxh": [240, 180], "channels": 3,
                       "l2_policy": "256 MB L2 flush (memset) is NOT used: 4 resident recordings rotate (200 MB > 126 MB L2)",
                       "parallelism": "independent recordings per GPU (no collective)"}

    @staticmethod
    def synth_recording(rng, n, W=240, H=180):
        """n synthetic N-Caltech101 records (5 bytes each: x, y, polarity bit + 23-bit big-endian timestamp)."""
        import numpy as np
        b = np.zeros((n, 5), dtype=np.uint8)
        b[:, 0] = rng.integers(0, W, n)
        b[:, 1] = rng.integers(0, H, n)
        t = np.sort(rng.integers(0, 1 << 23, n))
        b[:, 2] = ((t >> 16) & 0x7f) | (rng.integers(0, 2, n) << 7)
        b[:, 3] = (t >> 8) & 0xff
        b[:, 4] = t & 0xff
        return b.tobytes()

    def setup(self, torch, rank):
        import numpy as np
        from mem_b200 import _lib
        from mem_b200.process_data import RAW_NCALTECH101, histogram_raw
        self.torch, self._lib, self.fmt, self.histogram_raw = torch, _lib, RAW_NCALTECH101, histogram_raw
        self.host = [torch.frombuffer(bytearray(self.synth_recording(np.random.default_rng(10 * rank + k), self.n)),
                                      dtype=torch.uint8).pin_memory() for k in range(4)]
        self.dev = [h.cuda() for h in self.host]
        self.k, self.out = 0, None
        self.units_per_step = self.n
        self.h2d, self.d2h = self.n * 5, self.H * self.W * self.C

    def step_device(self):
        self.k += 1
        self.out = self.histogram_raw(self.dev[self.k % 4], self.fmt, self.H, self.W, channels=self.C, check=False)

    def step_e2e(self):
        self.k += 1
        dev = self.host[self.k % 4].to("cuda", non_blocking=True)
        return self.histogram_raw(dev, self.fmt, self.H, self.W, channels=self.C, check=True).cpu()

    def verify(self):
        import numpy as np
        from oracle import decode_ref                       # the checker, not the thing measured
        from oracle.histogram_ref import event_hist_ref
        n = min(self.n, 2_000_000)
        raw = self.host[0][:5 * n]
        got = self.histogram_raw(raw.cuda(), self.fmt, self.H, self.W, channels=self.C).cpu().numpy()
        return bool(np.array_equal(got, event_hist_ref(decode_ref.ncaltech101_np(raw.numpy().tobytes()), self.H, self.W)))

    def roofline(self, ms_per_step):
        hbm, _, _, how = measured_peaks()
        alg_bytes = 5.0 * self.n + self.C * self.H * self.W
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "hist_private<NCALTECH101> (records decoded inside the privatised rasteriser) + "
                                          "hist_private_finalize: whole step timed; at 5 B/event the step is bound by per-event "
                                          "instructions / shared-memory atomics and by the host call, not by HBM bytes",
                "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": None}

    def cpu_baseline(self):
        import numpy as np
        from oracle import decode_ref
        n = 200_000
        raw = self.synth_recording(np.random.default_rng(0), n)
        t0 = time.perf_counter()
        ev = decode_ref.ncaltech101_loop(raw)                # the reference's byte-by-byte Python decoder
        t1 = time.perf_counter()
        self.cpu_once(ev)                                    # + np.add.at rasteriser
        t2 = time.perf_counter()
        return {"value": round(n / (t2 - t0) / 1e9, 6), "unit": self.unit, "cores": 1, "kind": "port",
                "sample": f"{n} records: oracle/decode_ref.py ncaltech101_loop ({(t1 - t0):.2f} s) + oracle/histogram_ref.py "
                          f"({(t2 - t1) * 1e3:.0f} ms), 1 thread"}


# ----------------------------------------------------------------------------- event pipeline workload
class EventPipelineWorkload:
    """One training batch of raw streams -> model input: augment + rasterise + crop / hot-pixel filter / normalise.
    128 streams x 45000 raw events at 640x480 (N-ImageNet), 30000-event windows, 256x341 raster, 224x224 crops."""
    metric = "Gevents/s event pipeline (augment+rasterise+post)"
    unit = "Gevents/s"
    dtype = "f64->u8->f32"

    def __init__(self, args):
        self.B, self.n_raw, self.C = args.batch, 45000, 3
        self.config = {"workload": f"build_transformNPY(train) on the GPU: {self.B} streams x {self.n_raw} raw events (640x480), "
                                   "scale 256/480, 30000-event window, time/x flips, shift+cull, 256x341 raster, 224x224 crop, "
                                   "hot-pixel filter (10 sigma), normalise -> float32 [B,3,224,224]",
                       "batch": self.B, "raw_events_per_stream": self.n_raw, "window": 30000,
                       "l2_policy": "raw batch (184 MB) larger than the 126 MB L2; 4 resident batches rotate",
                       "parallelism": "independent batches per GPU (no collective)"}

    def make_stream(self, rng, n=None):
        import numpy as np
        n = n or self.n_raw
        ev = np.empty((n, 4), dtype=np.float64)
        seg = rng.integers(0, 12, n)
        x0, y0 = rng.uniform(0, 640, 12), rng.uniform(0, 480, 12)
        dx, dy = rng.uniform(-1, 1, 12), rng.uniform(-1, 1, 12)
        sgm = rng.uniform(0, 240, n)
        ev[:, 0] = np.clip(x0[seg] + dx[seg] * sgm + rng.normal(0, 1.5, n), 0, 639.99)
        ev[:, 1] = np.clip(y0[seg] + dy[seg] * sgm + rng.normal(0, 1.5, n), 0, 479.99)
        noise = rng.random(n) < 0.2
        ev[noise, 0] = rng.uniform(0, 639.99, int(noise.sum()))
        ev[noise, 1] = rng.uniform(0, 479.99, int(noise.sum()))
        ev[:, 2] = np.sort(rng.uniform(0, 3e5, n))
        ev[:, 3] = rng.integers(0, 2, n) * 2.0 - 1.0
        return ev

    def setup(self, torch, rank):
        import numpy as np
        from mem_b200 import _lib, event_pipeline as ep
        self.torch, self._lib, self.ep = torch, _lib, ep
        self.cfg = ep.PipelineConfig(is_train=True, normalize_events=True)
        self.pipe = ep.EventBatchPipeline(self.cfg, channels=self.C)
        rng = np.random.default_rng(100 + rank)
        base = [self.make_stream(rng) for _ in range(8)]
        self.host, self.dev = [], []
        for k in range(4):                      # 4 resident batches, streams drawn from 8 templates with a per-slot jitter
            ev = np.concatenate([base[(i + k) % 8] for i in range(self.B)], axis=0)
            ev[:, 0] = np.clip(ev[:, 0] + rng.uniform(-20, 20, len(ev)), 0, 639.99)
            h = torch.from_numpy(ev).pin_memory()
            self.host.append(h)
            self.dev.append(h.cuda())
        self.off_host = np.arange(self.B + 1, dtype=np.int64) * self.n_raw
        self.off_dev = torch.from_numpy(self.off_host).cuda()
        # device-resident arm: the per-sample augmentation records are inputs too (drawn once per resident batch)
        self.resident = []
        for k in range(4):
            aug, crop = ep.pack_params([ep.draw_params(self.n_raw, self.cfg) for _ in range(self.B)])
            self.resident.append((torch.from_numpy(aug.view(np.uint8).reshape(self.B, 64)).cuda(), torch.from_numpy(crop).cuda()))
        self.hw = self.cfg.raster_hw()
        self.out_host = None
        self.first_stream = base[0]
        self.k = 0
        self.units_per_step = self.B * self.cfg.slice_max_evs
        self.h2d, self.d2h = self.B * self.n_raw * 32, self.B * self.C * 224 * 224 * 4
        self.out = None

    def _run(self, events):
        lens = [self.n_raw] * self.B
        params = [self.ep.draw_params(n, self.cfg) for n in lens]     # host draws, reference generator order
        return self.pipe(events, self.off_dev, params=params)

    def step_device(self):
        self.k += 1
        aug, crop = self.resident[self.k % 4]
        self.out = self.ep.pipeline_fused(self.dev[self.k % 4], self.off_dev, aug, crop, self.hw[0], self.hw[1], (224, 224),
                                          self.C, hot_num_stds=10.0, normalize=True, check=False, out=self.out)

    def step_e2e(self):
        # public API from HOST buffers: generator draws, H2D of the raw batch, the kernel, D2H of the model input
        self.k += 1
        dev = self.host[self.k % 4].to("cuda", non_blocking=True)
        res = self._run(dev)
        if self.out_host is None:
            self.out_host = self.torch.empty(res.shape, dtype=res.dtype).pin_memory()
        self.out_host.copy_(res, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        return self.out_host

    def with_rand_aug(self, steps=50, warm=10):
        """Sub-record: the same batch with the reference scripts' default tail (ToUnit8 -> EventRandAugment(magnitude=20) ->
        ToFloat32, mem/datasets.py:655-658) as one more launch; operations drawn once per resident batch like the other
        per-sample records of the device-resident arm."""
        import contextlib
        import io
        from mem_b200 import transforms as T
        torch = self.torch
        with contextlib.redirect_stdout(io.StringIO()):
            aug = T.EventRandAugment(small=False, magnitude=20)
        torch.manual_seed(11)
        ops = [T.ops_to_device(aug.draw_batch(self.B, 224, 224), "cuda") for _ in range(4)]
        out2 = torch.empty(self.B, self.C, 224, 224, dtype=torch.float32, device="cuda")

        def step():
            self.step_device()
            T.apply_ops(self.out, ops[self.k % 4], out_float=True, out=out2)

        ms = _time_steps(torch, step, steps, warm)
        return {"value": round(self.units_per_step / (ms * 1e-3) / 1e9, 3), "unit": self.unit, "us_per_step": round(ms * 1e3, 2),
                "gpu_launches_per_step": 2,
                "what": "event_pipeline_fused + event_randaug (two drawn operations per sample), device-resident"}

    def verify(self):
        import random
        import numpy as np
        from oracle.event_pipeline_ref import PipelineCfg, pipeline_ref
        torch = self.torch
        ok = True
        for b in (0, 5):
            ev = self.host[0][b * self.n_raw:(b + 1) * self.n_raw].numpy()
            for g in (random, np.random, torch):
                (g.seed if g is not torch else g.manual_seed)(7 + b)
            want = pipeline_ref(ev, PipelineCfg(is_train=True, normalize_events=True)).numpy()
            for g in (random, np.random, torch):
                (g.seed if g is not torch else g.manual_seed)(7 + b)
            got = self.pipe([ev])[0].cpu().numpy()
            ok = ok and bool(np.array_equal(got, want))
        return ok

    def roofline(self, ms_per_step):
        hbm, _, _, how = measured_peaks()
        alg_bytes = 32.0 * self.units_per_step + self.d2h
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "event_pipeline_fused (one CTA per stream: augment + rasterise + crop + hot-pixel filter "
                                          "+ normalise in shared memory)",
                "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "algorithmic_bytes_per_launch": alg_bytes, "traffic": None}

    def cpu_baseline(self):
        import numpy as np
        from oracle.event_pipeline_ref import PipelineCfg, pipeline_ref
        cfg = PipelineCfg(is_train=True, normalize_events=True)
        ev = self.first_stream
        pipeline_ref(ev, cfg)
        reps, t0 = 0, time.perf_counter()
        while reps < 8 or time.perf_counter() - t0 < 4.0:
            pipeline_ref(ev, cfg)
            reps += 1
            if time.perf_counter() - t0 > 25:
                break
        dt = (time.perf_counter() - t0) / reps
        return {"value": round(cfg.slice_max_evs / dt / 1e9, 6), "unit": self.unit, "cores": 1, "kind": "port",
                "sample": f"{reps} samples of {self.n_raw} raw events through oracle/event_pipeline_ref.py "
                          f"(the reference's numpy/torch chain, 1 thread; {dt * 1e3:.1f} ms per sample)"}


def _ref_pipe_init(n_raw):
    global _REF_EV
    import numpy as np
    _REF_EV = EventPipelineWorkload.make_stream(None, np.random.default_rng(os.getpid()), n_raw)


def _ref_pipe_worker(k):
    import torch
    from oracle.event_pipeline_ref import PipelineCfg, pipeline_ref
    torch.set_num_threads(1)
    cfg = PipelineCfg(is_train=True, normalize_events=True)
    return float(sum(pipeline_ref(_REF_EV, cfg).sum() for _ in range(k)))


def run_reference_event_pipeline(args, wl):
    """Reference CPU arm: every host core runs the per-sample chain on its own resident stream (DataLoader workers)."""
    import multiprocessing as mp
    workers = max(1, min(os.cpu_count() or 1, 64))
    per_worker = 8
    steps, warm = args.steps or 5, args.warmup if args.warmup is not None else 1
    with mp.get_context("fork").Pool(workers, initializer=_ref_pipe_init, initargs=(wl.n_raw,)) as pool:
        for _ in range(max(1, warm)):
            pool.map(_ref_pipe_worker, [per_worker] * workers, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_ref_pipe_worker, [per_worker] * workers, chunksize=1)
        dt = time.perf_counter() - t0
    value = workers * per_worker * 30000 * steps / dt / 1e9
    return value, dt / steps * 1e3, workers, (f"{workers} worker processes (all host cores, capped at 64), each running the "
                                              f"per-sample chain on {per_worker} samples of {wl.n_raw} raw events per step")


_REF_EV = None


def _ref_hist_init(n, H, W):
    global _REF_EV
    import numpy as np
    rng = np.random.default_rng(os.getpid())
    ev = np.empty((n, 4), dtype=np.float64)
    ev[:, 0] = rng.integers(0, W, n); ev[:, 1] = rng.integers(0, H, n)
    ev[:, 2] = np.sort(rng.uniform(0, 3e5, n)); ev[:, 3] = rng.integers(0, 2, n) * 2.0 - 1.0
    _REF_EV = ev


def _ref_hist_worker(a):
    H, W = a
    from oracle.histogram_ref import event_hist_ref
    return int(event_hist_ref(_REF_EV, H, W)[0, 0, 0])


def run_reference_histogram(args, wl):
    """Reference CPU arm: every host core rasterises its own resident stream (how the reference's
    DataLoader workers parallelise EventArrToImg, run_mem_pretraining.py:333-339)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    n_each = 1_000_000
    steps, warm = args.steps or 10, args.warmup if args.warmup is not None else 3
    with mp.get_context("fork").Pool(workers, initializer=_ref_hist_init, initargs=(n_each, wl.H, wl.W)) as pool:
        for _ in range(max(1, warm)):
            pool.map(_ref_hist_worker, [(wl.H, wl.W)] * workers, chunksize=1)
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(_ref_hist_worker, [(wl.H, wl.W)] * workers, chunksize=1)
        dt = time.perf_counter() - t0
    value = workers * n_each * steps / dt / 1e9
    return value, dt / steps * 1e3, workers, (f"{workers} worker processes (all host cores, capped at 64), each "
                                              f"rasterising a resident {n_each}-event stream per step")


def _time_steps(torch, fn, steps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def histogram_record(args):
    """`Gevents/s histogram rasterise` (BASELINE.json's second metric) at the top size of config 2 -- 10 M events on the
    640x480 N-ImageNet sensor (`streams`) and on the 240x180 N-Caltech101 sensor (`streams_240x180`) -- for the spatial
    distributions of SURVEY.md 8(d), each checked against the oracle."""
    import torch
    rec = {"metric": HistogramWorkload.metric, "unit": HistogramWorkload.unit, "dtype": HistogramWorkload.dtype,
           "config": {"workload": "event->histogram rasterise, 10000000 events, 640x480 sensor, float64[N,4] rows -> uint8[H,W,3]",
                      "l2_policy": "input (320 MB) larger than the 126 MB L2", "steps": 20, "warmup": 3}, "streams": {}}
    for kind in ("uniform", "edge8", "edge50", "hot"):
        wl = HistogramWorkload(argparse.Namespace(sensor="640x480", events=10_000_000, dist=kind))
        wl.setup(torch, 0)
        ok = wl.verify()
        l0 = wl._lib.launch_count()
        ms = _time_steps(torch, wl.step_device, 20, 3)
        launches = (wl._lib.launch_count() - l0) // 23
        roof = wl.roofline(ms)
        rec["streams"][kind] = {"value": round(wl.n / (ms * 1e-3) / 1e9, 2), "us_per_step": round(ms * 1e3, 2), "gpu_launches_per_step": int(launches),
                                "parity_vs_oracle": ok, "roofline": roof}
        if kind == "uniform":
            ms_e2e = _time_steps(torch, wl.step_e2e, 5, 2)
            rec["e2e"] = {"value": round(wl.n / (ms_e2e * 1e-3) / 1e9, 4), "unit": wl.unit, "h2d_bytes_per_step": wl.h2d,
                          "d2h_bytes_per_step": wl.d2h, "ms_per_step": round(ms_e2e, 3), "api": "process_data.histogram (pinned host rows in, host image out)"}
            if not args.no_cpu_baseline:
                rec["cpu_baseline"] = wl.cpu_baseline()
        del wl
        torch.cuda.empty_cache()
    vals = [v["value"] for v in rec["streams"].values()]
    rec["value"] = rec["streams"]["uniform"]["value"]
    rec["worst_stream_value"] = min(vals)
    rec["roofline"] = rec["streams"]["uniform"]["roofline"]
    # the other sensor of config 2 (N-Caltech101, 240x180): the whole sensor fits one SM's shared memory (strategy PRIVATE)
    rec["streams_240x180"] = {}
    for kind in ("uniform", "edge8", "hot"):
        wl = HistogramWorkload(argparse.Namespace(sensor="240x180", events=10_000_000, dist=kind))
        wl.setup(torch, 0)
        ok = wl.verify()
        ms = _time_steps(torch, wl.step_device, 20, 3)
        rec["streams_240x180"][kind] = {"value": round(wl.n / (ms * 1e-3) / 1e9, 2), "us_per_step": round(ms * 1e3, 2),
                                        "parity_vs_oracle": ok, "roofline": wl.roofline(ms)}
        del wl
        torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------- main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    workload = args.workload
    if workload == "auto":
        workload = "pretrain"
    if workload == "pretrain":
        from benchmarks import pretrain
        line = pretrain.main(args, rank, local_rank, world, ClockSampler, measured_peaks)
        if line is not None:
            # the second metric BASELINE.json names rides on the same line (single-GPU runs: rasterisation has no collective)
            if args.impl != "reference" and world == 1 and not args.no_histogram:
                line["histogram"] = histogram_record(args)
            print(json.dumps(line))
        return
    if workload == "finetune":
        from benchmarks import finetune
        return finetune.main(args, rank, local_rank, world, ClockSampler, measured_peaks)

    wl = {"event_pipeline": EventPipelineWorkload, "raw_histogram": RawHistogramWorkload}.get(workload, HistogramWorkload)(args)

    if args.impl == "reference":
        if rank != 0:
            return
        if workload == "raw_histogram":
            # the reference decodes byte by byte in one Python process per recording: a single-core sample is the arm
            cb = wl.cpu_baseline()
            records = 200_000                                 # what RawHistogramWorkload.cpu_baseline decodes
            print(json.dumps({"impl": "reference", "metric": wl.metric, "value": cb["value"], "unit": wl.unit, "n_gpus": args.gpus,
                              "steps": 1, "warmup": 0, "ms_per_step": round(records / cb["value"] / 1e6, 3), "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config,
                              "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": wl.unit, "h2d_bytes_per_step": 0,
                                                          "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
            return
        value, ms, workers, sample = (run_reference_event_pipeline if workload == "event_pipeline" else run_reference_histogram)(args, wl)
        line = {"impl": "reference", "metric": wl.metric, "value": round(value, 6), "unit": wl.unit,
                "n_gpus": args.gpus, "steps": args.steps or 10, "warmup": args.warmup if args.warmup is not None else 3,
                "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": wl.dtype, "data": "synthetic", "config": wl.config,
                "cpu_baseline": {"value": round(value, 6), "unit": wl.unit, "cores": workers, "kind": "port", "sample": sample},
                "e2e": {"value": round(value, 6), "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = args.steps or 50
    warm = args.warmup if args.warmup is not None else 10
    warm = max(warm, 3)

    wl.setup(torch, rank)
    ok = wl.verify()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(warm):
        wl.step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = wl._lib.launch_count()
    ms_total = timed(wl.step_device, steps)
    launches = wl._lib.launch_count() - l0
    # end-to-end through the public API with host buffers
    for _ in range(3):
        wl.step_e2e()
    e2e_steps = max(3, min(steps, 10))
    ms_e2e = timed(wl.step_e2e, e2e_steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / steps
    value = wl.units_per_step * world / (ms_step * 1e-3) / 1e9
    e2e_val = wl.units_per_step * world / (ms_e2e / e2e_steps * 1e-3) / 1e9
    line = {"metric": wl.metric, "value": round(value, 3), "unit": wl.unit, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config,
            "e2e": {"value": round(e2e_val, 4), "unit": wl.unit, "h2d_bytes_per_step": wl.h2d,
                    "d2h_bytes_per_step": wl.d2h, "ms_per_step": round(ms_e2e / e2e_steps, 3)},
            "gpu_launches": int(launches), "parity_vs_oracle": ok, "clocks": clocks,
            "roofline": wl.roofline(ms_step)}
    if hasattr(wl, "with_rand_aug") and world == 1:
        line["with_rand_aug"] = wl.with_rand_aug()
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = wl.cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout_main():
    """Library chatter (e.g. NCCL's version banner) goes to stderr; stdout carries the JSON line only."""
    import builtins
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    orig_print = builtins.print

    def json_print(*a, **k):
        if len(a) == 1 and isinstance(a[0], str) and a[0].startswith("{") and "file" not in k:
            out.write(a[0] + "\n")
            out.flush()
        else:
            orig_print(*a, **k)
    builtins.print = json_print
    try:
        main()
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        out.flush()


if __name__ == "__main__":
    _quiet_stdout_main()
