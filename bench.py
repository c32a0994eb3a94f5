#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Workload (default, BASELINE.json configs[1]): batched 4096-point radix-4 complex FFT, fp32, 65536 frames
per GPU, in place, frame-major interleaved complex.  One step = one pass of the transform over the whole
batch (4 GiB of algorithmic HBM traffic).  Steps alternate forward / reverse (the reverse carries the
1/N scale, same kernel, same bytes) so the data stays bounded and the run checks itself: after an even
number of steps the batch must equal the input again.

Prints ONE JSON line.  `value` = whole-job Msamples/s with the batch resident in HBM; `e2e` = the same
through the public host-buffer API (pinned host memory, H2D + D2H inside the timed region);
`roofline` = algorithmic bytes / CUDA-event kernel time against the measured HBM copy peak;
`cpu_baseline` = the reference's own CPU implementation (oracle/_ref) on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent

WORKLOADS = {
    # name: (kind, params)
    "fft4096_f32": dict(kind="fft", n=4096, frames=65536, precision="f32", bytes_per_sample=16),
    "fft4096_f64": dict(kind="fft", n=4096, frames=65536, precision="f64", bytes_per_sample=32),
    "fft65536_f32": dict(kind="fft", n=65536, frames=4096, precision="f32", bytes_per_sample=16),
    "fft1024_f32": dict(kind="fft", n=1024, frames=262144, precision="f32", bytes_per_sample=16),
    "fft1024_f64": dict(kind="fft", n=1024, frames=262144, precision="f64", bytes_per_sample=32),
    "fft256_f64": dict(kind="fft", n=256, frames=1 << 20, precision="f64", bytes_per_sample=32),
    "fft2048_f64": dict(kind="fft", n=2048, frames=131072, precision="f64", bytes_per_sample=32),
    "fft256_f32": dict(kind="fft", n=256, frames=1 << 20, precision="f32", bytes_per_sample=16),
    "fft64_f32": dict(kind="fft", n=64, frames=1 << 22, precision="f32", bytes_per_sample=16),
    "fft32768_f32": dict(kind="fft", n=32768, frames=8192, precision="f32", bytes_per_sample=16),
    # real frames in (4 B/sample), complex spectra out (8 B/sample), out of place: config 5's transform on its own
    "fftreal65536_f32": dict(kind="fft", n=65536, frames=4096, precision="f32", bytes_per_sample=12, real_input=True),
    # real frames in, half spectra out (sdsp_b200_fft_exec_r2c): 4 + 4 bytes per real sample
    "fftr2c4096_f32": dict(kind="fft", n=4096, frames=131072, precision="f32", bytes_per_sample=8, real_input=True, half=True),
    "fftr2c4096_f64": dict(kind="fft", n=4096, frames=65536, precision="f64", bytes_per_sample=16, real_input=True, half=True),
    "fftr2c65536_f32": dict(kind="fft", n=65536, frames=4096, precision="f32", bytes_per_sample=8, real_input=True, half=True),
    # half spectra in, real frames out (sdsp_b200_fft_exec_c2r, 1/n included)
    "fftc2r4096_f32": dict(kind="fft", n=4096, frames=131072, precision="f32", bytes_per_sample=8, real_input=True, half=True, back=True),
    "fft131072_f32": dict(kind="fft", n=131072, frames=2048, precision="f32", bytes_per_sample=16),
    "fft262144_f32": dict(kind="fft", n=262144, frames=1024, precision="f32", bytes_per_sample=16),
    "fft16384_f32": dict(kind="fft", n=16384, frames=16384, precision="f32", bytes_per_sample=16),
    "fft8192_f32": dict(kind="fft", n=8192, frames=32768, precision="f32", bytes_per_sample=16),
    "fft2048_f32": dict(kind="fft", n=2048, frames=131072, precision="f32", bytes_per_sample=16),
    "iir16384_f32": dict(kind="iir", channels=16384, samples=1 << 20, precision="f32", sections=4, bytes_per_sample=8),
    "iir4096_f32": dict(kind="iir", channels=4096, samples=1 << 22, precision="f32", sections=4, bytes_per_sample=8),
    "iir4096_f32_scan": dict(kind="iir", channels=4096, samples=1 << 22, precision="f32", sections=4, bytes_per_sample=8, path="scan"),
    "iirscan_f64": dict(kind="iir", channels=1, samples=1 << 30, precision="f64", sections=4, bytes_per_sample=16, path="scan"),
    "iirscan_f32": dict(kind="iir", channels=1, samples=1 << 30, precision="f32", sections=4, bytes_per_sample=8, path="scan"),
    "iirscan_f64_lookback": dict(kind="iir", channels=1, samples=1 << 30, precision="f64", sections=4, bytes_per_sample=16, path="lookback"),
    "iir4096_f32_lookback": dict(kind="iir", channels=4096, samples=1 << 22, precision="f32", sections=4, bytes_per_sample=8, path="lookback"),
    # BASELINE config 5, one GPU's share at 8 GPUs: 512 channels x 64 frames of 65536 real samples (32768 frames):
    # IIR along every channel (time-split path), then every frame transformed (real in, spectrum out)
    "pipeline65536_f32": dict(kind="pipeline", channels=512, frames_per_channel=64, n=65536, precision="f32", sections=4,
                              bytes_per_sample=20),
    # BASELINE config 5 as stated: 262144 frames of 65536 points = 4096 channels x 64 frames, IIR then real-input FFT, the
    # WHOLE job divided over the ranks (strong scaling): rank r filters its 4096/N channels in one call and transforms
    # them in waves of 32768 frames (one 16 GiB spectrum buffer)
    "pipeline_cfg5_f32": dict(kind="pipeline", channels=4096, frames_per_channel=64, n=65536, precision="f32", sections=4,
                              bytes_per_sample=20, strong=True, wave_frames=32768),
    # the same job with half spectra out (bins 0 .. 32768 of every frame; the rest is the conjugate mirror): 8 + 4 + 4 bytes per sample
    "pipeline_cfg5_r2c_f32": dict(kind="pipeline", channels=4096, frames_per_channel=64, n=65536, precision="f32", sections=4,
                                  bytes_per_sample=16, strong=True, wave_frames=65536, half=True),
    "iir16384_f32_scan": dict(kind="iir", channels=16384, samples=1 << 20, precision="f32", sections=4, bytes_per_sample=8, path="scan"),
    # same bank, channel pitch not a power of two (2^20 + 8256 samples): separates DRAM channel effects from kernel effects
    "iir16384_f32_pitch": dict(kind="iir", channels=16384, samples=1 << 20, pitch=(1 << 20) + 8256, precision="f32", sections=4,
                               bytes_per_sample=8),
    # 18944 channels = 592 warps = exactly four row-warps on every one of the 148 SMs (16384 channels leave 3 or 4 per SM)
    "iir18944_f32": dict(kind="iir", channels=18944, samples=1 << 20, precision="f32", sections=4, bytes_per_sample=8),
    "iir16384_f64": dict(kind="iir", channels=16384, samples=1 << 19, precision="f64", sections=4, bytes_per_sample=16),
}


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch at the workload's full size, from the
# committed `ncu --set full` summaries (profiles/); None where no capture at that size exists
NCU_TRAFFIC = {
    "fft4096_f32": (4.26326e9, "profiles/r02_ncu_fft4096_f32_v1.txt"),
    "fft4096_f64": (8.683e9, "profiles/r01_ncu_fft4096_f64_v2.txt"),
    "fft65536_f32": (4.39971e9, "profiles/r02_fft_lag_traffic.txt (lead 768, bench size)"),
    "iir16384_f32": (1.37387e11, "profiles/r02_ncu_iir16384_f32_delta_v1.txt"),
    "iir4096_f32_scan": (1.37408e11, "profiles/r01_ncu_iir4096_f32_split_v1.txt (main pass)"),
    "iirscan_f64": (1.7126e10, "profiles/r01_launches_iir_split_v1.txt (main pass)"),
    # captures taken on a quarter / an eighth of the bench's frames (ncu replays every launch), scaled to the bench's frame count
    "fft8192_f32": (4 * 1.03341e9, "profiles/r02_ncu_fft8192_f32_radix32_v1.txt (8192 frames: 1.033 GB for 1.074 GB algorithmic; x 4)"),
    "fft16384_f32": (4 * 1.02115e9, "profiles/r02_ncu_fft16384_f32_radix32_v1.txt (4096 frames: 1.021 GB for 1.074 GB algorithmic; x 4)"),
    "fftr2c4096_f32": (8 * 4.92013e8, "profiles/r02_ncu_fftr2c4096_f32_v1.txt (16384 frames: 0.492 GB for 0.537 GB algorithmic; x 8)"),
    "fftreal65536_f32": (3.84578e9, "profiles/r02_fft_lag_traffic.txt (lag 96, bench size: 3.85 GB for 3.22 GB algorithmic -- a fifth of the ring spills, accepted for speed)"),
    "fftr2c65536_f32": (2.14801e9, "profiles/r02_fft_lag_traffic.txt (lag 96, bench size)"),
}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {
        0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
        0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
def dist_setup(n_gpus: int):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if n_gpus != world:
        if rank == 0:
            print(f"warning: --gpus {n_gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    return rank, local, world


_ORIGINAL_AFFINITY = None


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host memory is allocated or
    touched: the end-to-end leg moves 4 GiB per step over PCIe per rank, and with 8 ranks the staging buffers must
    not all land on one socket."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            global _ORIGINAL_AFFINITY
            _ORIGINAL_AFFINITY = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass  # no NVML / not permitted: keep the inherited affinity


def barrier(world):
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value: float, world: int) -> float:
    import torch
    import torch.distributed as dist

    if world == 1:
        return value
    t = torch.tensor([value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------------
class FftWorkload:
    def __init__(self, spec, device):
        import torch

        import simpledsp_b200 as S
        from simpledsp_b200 import _capi as K

        self.spec, self.S, self.K, self.torch = spec, S, K, torch
        self.n, self.frames = spec["n"], spec["frames"]
        self.prec = K.F32 if spec["precision"] == "f32" else K.F64
        self.rdtype = torch.float32 if self.prec == K.F32 else torch.float64
        self.radix = 4 if (self.n.bit_length() - 1) % 2 == 0 else 2
        self.fwd = S.FftPlan(self.n, self.radix, self.prec, K.FORWARD, device)
        self.inv = S.FftPlan(self.n, self.radix, self.prec, K.REVERSE, device)
        g = torch.Generator(device="cuda").manual_seed(1234 + device)
        self.half = bool(spec.get("half"))
        self.back = bool(spec.get("back"))
        if self.back:  # half spectra in (of real signals: bins 0 and n/2 real)
            self.data = torch.randn(self.frames, self.n // 2 + 1, 2, device="cuda", generator=g, dtype=self.rdtype)
            self.data[:, 0, 1] = 0
            self.data[:, -1, 1] = 0
        elif self.half:  # half spectra out: frames of n/2 + 1 bins
            self.data = torch.zeros(self.frames, self.n // 2 + 1, 2, device="cuda", dtype=self.rdtype)
        else:
            self.data = torch.randn(self.frames, self.n, 2, device="cuda", generator=g, dtype=self.rdtype)
        self.check_idx = torch.linspace(0, self.frames - 1, 32).long().cuda()
        self.orig = self.data[self.check_idx].clone()
        self.samples_per_step = self.frames * self.n
        self.step_no = 0
        self.stream = torch.cuda.current_stream().cuda_stream
        self.real_input = bool(spec.get("real_input"))
        if self.real_input:
            self.real = torch.randn(self.frames, self.n, device="cuda", generator=g, dtype=self.rdtype)

    def describe(self):
        if self.half:
            return ("half spectra in, real frames out" if self.back else "real frames in, half spectra (n/2 + 1 bins) out") + ": the n/2-point configuration of: " + self.S.FftPlan(
                self.n // 2, 2, self.prec, self.K.FORWARD, self.fwd.device).describe()
        return ("real frames in, spectra out: " if self.real_input else "") + self.fwd.describe()

    def launches_per_step(self):
        return self.fwd.launches(self.frames)

    def step(self):
        if self.back:
            self.inv.exec_c2r_ptr(self.data.data_ptr(), self.real.data_ptr(), self.frames, self.K.PTR_DEVICE, self.stream)
            self.step_no += 1
            return
        if self.half:
            self.fwd.exec_r2c_ptr(self.real.data_ptr(), self.data.data_ptr(), self.frames, self.K.PTR_DEVICE, self.stream)
            self.step_no += 1
            return
        if self.real_input:
            self.fwd.exec_real_ptr(self.real.data_ptr(), self.data.data_ptr(), self.frames, self.K.PTR_DEVICE, self.stream)
            self.step_no += 1
            return
        plan = self.fwd if self.step_no % 2 == 0 else self.inv
        plan.exec_ptr(self.data.data_ptr(), self.frames, self.K.PTR_DEVICE, self.stream)
        self.step_no += 1

    def self_check(self):
        """After an even number of steps the batch is the input again (forward then reverse).  Real input: sampled spectra
        against torch.fft of the same frames."""
        if self.back:
            torch = self.torch
            torch.cuda.synchronize()
            ref = torch.fft.irfft(torch.view_as_complex(self.data[self.check_idx]).to(torch.complex128), n=self.n)
            got = self.real[self.check_idx].double()
            return float(((got - ref).pow(2).sum(dim=1).sqrt() / ref.pow(2).sum(dim=1).sqrt()).max())
        if self.real_input:
            torch = self.torch
            torch.cuda.synchronize()
            ref = (torch.fft.rfft if self.half else torch.fft.fft)(self.real[self.check_idx].double())
            got = torch.view_as_complex(self.data[self.check_idx]).to(torch.complex128)
            return float(((got - ref).abs().pow(2).sum(dim=1).sqrt() / ref.abs().pow(2).sum(dim=1).sqrt()).max())
        if self.step_no % 2:
            self.step()
        self.torch.cuda.synchronize()
        now = self.data[self.check_idx]
        err = (now - self.orig).pow(2).sum(dim=(1, 2)).sqrt() / self.orig.pow(2).sum(dim=(1, 2)).sqrt()
        return float(err.max())

    # ---- end to end through the host-buffer API
    def e2e_prepare(self):
        import simpledsp_b200._capi as K

        if self.half:
            return self._e2e_prepare_half()
        nbytes = self.frames * self.n * 2 * (4 if self.prec == K.F32 else 8)
        ptr = C.c_void_p()
        K.check(K.lib().sdsp_b200_host_alloc(C.byref(ptr), nbytes))
        self._pinned = ptr
        cdt = np.complex64 if self.prec == K.F32 else np.complex128
        buf = (C.c_char * nbytes).from_address(ptr.value)
        self.host = np.frombuffer(buf, dtype=cdt).reshape(self.frames, self.n)
        rng = np.random.default_rng(99)
        blk = (rng.standard_normal((1024, self.n)) + 1j * rng.standard_normal((1024, self.n))).astype(cdt)
        for i in range(0, self.frames, 1024):
            self.host[i:i + 1024] = blk[: min(1024, self.frames - i)]
        self.host_first = self.host[:4].copy()
        self.e2e_steps_done = 0
        return nbytes, nbytes

    def _e2e_prepare_half(self):
        """Real frames in pinned host memory -> half spectra in pinned host memory through FftPlan.half_spectrum."""
        import simpledsp_b200._capi as K

        es = 4 if self.prec == K.F32 else 8
        in_bytes, out_bytes = self.frames * self.n * es, self.frames * (self.n // 2 + 1) * 2 * es
        self._pinned, self._pinned_out = C.c_void_p(), C.c_void_p()
        K.check(K.lib().sdsp_b200_host_alloc(C.byref(self._pinned), in_bytes))
        K.check(K.lib().sdsp_b200_host_alloc(C.byref(self._pinned_out), out_bytes))
        rdt, cdt = (np.float32, np.complex64) if self.prec == K.F32 else (np.float64, np.complex128)
        self.host = np.frombuffer((C.c_char * in_bytes).from_address(self._pinned.value), dtype=rdt).reshape(self.frames, self.n)
        self.host_out = np.frombuffer((C.c_char * out_bytes).from_address(self._pinned_out.value), dtype=cdt).reshape(self.frames, self.n // 2 + 1)
        blk = np.random.default_rng(99).standard_normal((1024, self.n)).astype(rdt)
        for i in range(0, self.frames, 1024):
            self.host[i:i + 1024] = blk[: min(1024, self.frames - i)]
        self.e2e_steps_done = 0
        return in_bytes, out_bytes

    def e2e_step(self):
        if self.half:
            self.fwd.half_spectrum(self.host, out=self.host_out)
            self.e2e_steps_done += 1
            return
        plan = self.fwd if self.e2e_steps_done % 2 == 0 else self.inv
        plan(self.host)  # public API on a host array: H2D, transform, D2H, synchronous
        self.e2e_steps_done += 1

    def e2e_check(self):
        if self.half:
            ref = np.fft.rfft(self.host[-4:].astype(np.float64), axis=1)
            return float(np.linalg.norm(self.host_out[-4:] - ref) / np.linalg.norm(ref))
        if self.e2e_steps_done % 2:
            self.e2e_step()
        err = np.linalg.norm(self.host[:4] - self.host_first) / np.linalg.norm(self.host_first)
        return float(err)

    def e2e_release(self):
        import simpledsp_b200._capi as K

        self.host = None
        K.lib().sdsp_b200_host_free(self._pinned)
        if self.half:
            self.host_out = None
            K.lib().sdsp_b200_host_free(self._pinned_out)


class IirWorkload:
    def __init__(self, spec, device):
        import torch

        import simpledsp_b200 as S
        from simpledsp_b200 import _capi as K

        self.spec, self.S, self.K, self.torch = spec, S, K, torch
        self.ch, self.n, self.m = spec["channels"], spec["samples"], spec["sections"]
        self.prec = K.F32 if spec["precision"] == "f32" else K.F64
        self.rdtype = torch.float32 if self.prec == K.F32 else torch.float64
        self.path = {"scan": K.IIR_SCAN, "lookback": K.IIR_SCAN_LOOKBACK}.get(spec.get("path"), K.IIR_AUTO)
        self.bank = S.IirBank(self.m, self.ch, self.prec, K.NUM_GENERIC, device)
        fs = 100e3
        ftype = np.where(np.arange(self.ch) % 2 == 0, K.LOW_PASS, K.HIGH_PASS)
        f0 = np.geomspace(1e3, 20e3, self.ch) if self.ch > 1 else np.array([10e3])
        uniq = {}
        gains, bs, as_ = np.zeros(self.ch), np.zeros((self.ch, self.m, 3)), np.zeros((self.ch, self.m, 3))
        for c in range(self.ch):
            key = (int(ftype[c]), float(f0[c]))
            if key not in uniq:
                uniq[key] = S.design(key[0], self.m, key[1], fs)
            gains[c], bs[c], as_[c] = uniq[key]
        self.bank.set_coeffs(gains, bs, as_)
        g = torch.Generator(device="cuda").manual_seed(1234 + device)
        self.pitch = spec.get("pitch", self.n)
        self.data = torch.empty(self.ch, self.pitch, device="cuda", dtype=self.rdtype)
        rows = max(1, (1 << 28) // self.pitch)
        for lo in range(0, self.ch, rows):
            self.data[lo:lo + rows].normal_(generator=g)
        self.samples_per_step = self.ch * self.n
        self.stream = torch.cuda.current_stream().cuda_stream

    def describe(self):
        return self.bank.describe(self.n, self.pitch, self.path)

    def launches_per_step(self):
        return 1  # the bandwidth-bound pass: all the algorithmic bytes go through one launch

    def gpu_launches_per_step(self):
        # time-split: gather, row pass, carry, correction pass (leftover rounds, if any, not counted); else one kernel
        return 4 if "time-split" in self.describe() else 1

    def step(self):
        self.bank.process_ptr(self.data.data_ptr(), self.n, self.pitch, self.K.PTR_DEVICE, self.path, self.stream)

    def self_check(self):
        self.torch.cuda.synchronize()
        return float(self.torch.isfinite(self.data[:: max(1, self.ch // 64), :4096]).all().item() == 0)

    # ---- end to end through the host-buffer API: a bounded stretch of the same bank (E2E_SAMPLES per channel) in pinned
    # host memory, staged in chunks of time by sdsp_b200_iir_bank_process(PTR_HOST)
    E2E_BYTES = 4 << 30

    def e2e_prepare(self):
        K = self.K
        es = 4 if self.prec == K.F32 else 8
        self.e2e_n = min(self.n, max(4096, self.E2E_BYTES // (self.ch * es)))
        nbytes = self.ch * self.e2e_n * es
        ptr = C.c_void_p()
        K.check(K.lib().sdsp_b200_host_alloc(C.byref(ptr), nbytes))
        self._pinned = ptr
        buf = (C.c_char * nbytes).from_address(ptr.value)
        self.host = np.frombuffer(buf, dtype=np.float32 if self.prec == K.F32 else np.float64).reshape(self.ch, self.e2e_n)
        rng = np.random.default_rng(99)
        blk = rng.standard_normal((min(self.ch, 256), self.e2e_n)).astype(self.host.dtype)
        for lo in range(0, self.ch, blk.shape[0]):
            self.host[lo:lo + blk.shape[0]] = blk[: min(blk.shape[0], self.ch - lo)]
        self.e2e_samples_per_step = self.ch * self.e2e_n
        self.bank.reset_state()
        return nbytes, nbytes

    def e2e_step(self):
        self.bank.process_ptr(self.host.ctypes.data, self.e2e_n, self.e2e_n, self.K.PTR_HOST, self.path, None)

    def e2e_check(self):
        return float(np.isfinite(self.host[:: max(1, self.ch // 64), :4096]).all() == 0)

    def e2e_release(self):
        self.host = None
        self.K.lib().sdsp_b200_host_free(self._pinned)


class PipelineWorkload:
    """IIR bank over [channels][frames*n] real samples in place, then a real-input FFT of every n-sample frame into
    a spectrum buffer.  Algorithmic bytes per sample: 8 (filter, read + write) + 4 + 8 (transform, real in, complex out)."""

    def __init__(self, spec, device, world=1):
        import torch

        import simpledsp_b200 as S
        from simpledsp_b200 import _capi as K

        self.spec, self.S, self.K, self.torch = spec, S, K, torch
        self.ch, self.fpc, self.n, self.m = spec["channels"], spec["frames_per_channel"], spec["n"], spec["sections"]
        if spec.get("strong"):  # the stated job divided over the ranks
            assert self.ch % world == 0
            self.ch //= world
        self.prec = K.F32
        self.len = self.fpc * self.n
        self.bank = S.IirBank(self.m, self.ch, self.prec, K.NUM_GENERIC, device)
        g_, b_, a_ = S.design(K.LOW_PASS, self.m, 10e3, 100e3)
        self.bank.set_coeffs(np.full(self.ch, g_), np.tile(b_, (self.ch, 1, 1)), np.tile(a_, (self.ch, 1, 1)))
        self.plan = S.FftPlan(self.n, 4, self.prec, K.FORWARD, device)
        g = torch.Generator(device="cuda").manual_seed(1234 + device)
        self.signal = torch.empty(self.ch, self.len, device="cuda", dtype=torch.float32)
        rows = max(1, (1 << 28) // self.len)
        for lo in range(0, self.ch, rows):
            self.signal[lo:lo + rows].normal_(generator=g)
        self.frames = self.ch * self.fpc
        self.wave = min(self.frames, spec.get("wave_frames", self.frames))
        self.half = bool(spec.get("half"))
        self.bins = self.n // 2 + 1 if self.half else self.n
        self.spectra = torch.empty(self.wave, self.bins, device="cuda", dtype=torch.complex64)
        self.samples_per_step = self.ch * self.len
        self.stream = torch.cuda.current_stream().cuda_stream

    def describe(self):
        return (self.bank.describe(self.len, self.len, self.K.IIR_SCAN) + f" | then real-input, {self.frames} frames in waves of {self.wave}"
                + (", half spectra out: " if self.half else ": ") + self.plan.describe())

    def launches_per_step(self):
        return 1

    def gpu_launches_per_step(self):
        # filter: gather, row pass, carry, correction pass; transform: one persistent kernel per wave
        return 4 + (self.frames + self.wave - 1) // self.wave

    def step(self):
        K = self.K
        self.bank.process_ptr(self.signal.data_ptr(), self.len, self.len, K.PTR_DEVICE, K.IIR_SCAN, self.stream)
        for lo in range(0, self.frames, self.wave):
            cnt = min(self.wave, self.frames - lo)
            (self.plan.exec_r2c_ptr if self.half else self.plan.exec_real_ptr)(
                self.signal.data_ptr() + lo * self.n * 4, self.spectra.data_ptr(), cnt, K.PTR_DEVICE, self.stream)

    def self_check(self):
        """A sampled frame of the spectrum buffer equals the transform of the filtered signal it was made from."""
        torch = self.torch
        torch.cuda.synchronize()
        worst = 0.0
        last_wave = (self.frames - 1) // self.wave * self.wave  # the spectrum buffer holds the last wave
        for fr in (last_wave, last_wave + (self.frames - last_wave) // 2, self.frames - 1):
            x = self.signal.view(-1)[fr * self.n:(fr + 1) * self.n].double()
            ref = (torch.fft.rfft if self.half else torch.fft.fft)(x)
            got = self.spectra[fr - last_wave].to(torch.complex128)
            worst = max(worst, float((got - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt().clamp_min(1e-300)))
        return worst

    def e2e_prepare(self):
        return None

    def e2e_release(self):
        pass


# --------------------------------------------------------------------------------------------------
def config_of(name, spec):
    """The `config` object of the JSON line -- built the same way by both arms (b200 and --impl reference)."""
    cfg = {"workload": name, **{k: v for k, v in spec.items() if k != "kind"}}
    cfg["per_gpu"] = not spec.get("strong", False)
    cfg["in_place"] = True
    cfg["l2"] = "no flush needed: the working set of one step is far larger than the 126 MB L2"
    return cfg


def cpu_reference_rate(spec, threads: int, target_seconds: float = 6.0):
    """The reference's own CPU implementation on a bounded sample of the workload.
    -> (Msamples/s, kind, cores, sample description, single-thread Msamples/s)"""
    from oracle import oracle as O

    if spec["kind"] == "pipeline":  # filter then transform: per-sample times add
        a = cpu_reference_rate(dict(kind="iir", channels=spec["channels"], samples=spec["frames_per_channel"] * spec["n"],
                                    sections=spec["sections"]), threads, target_seconds / 2)
        b = cpu_reference_rate(dict(kind="fft", n=spec["n"]), threads, target_seconds / 2)
        rate = 1.0 / (1.0 / a[0] + 1.0 / b[0])
        single = 1.0 / (1.0 / a[4] + 1.0 / b[4])
        kind = "reference" if a[1] == b[1] == "reference" else "port"
        return rate, kind, max(a[2], b[2]), f"filter: {a[3]} | transform: {b[3]}", single
    kind = "reference" if O.have_ref() else "port"
    rng = np.random.default_rng(1234)
    if spec["kind"] == "fft":
        n = spec["n"]
        radix = 4 if (n.bit_length() - 1) % 2 == 0 else 2
        # the compiled reference covers every power of two up to 4096 and, in the _big library, 16384 and 65536
        compiled = kind == "reference" and (n <= 4096 or (n in (16384, 65536) and os.path.exists(O.REF_BIG_SO)))
        impl = "reference" if compiled else "port"
        kind = impl
        probe = rng.standard_normal((16 if n > 4096 else 64, n)) + 1j * rng.standard_normal((16 if n > 4096 else 64, n))
        O.fft(probe[:2], radix, False, impl)  # builds / pages in the tables
        t0 = time.perf_counter()
        O.fft(probe, radix, False, impl, threads=1)
        one = (time.perf_counter() - t0) / probe.shape[0]
        use_threads = threads if impl == "reference" else 1
        frames = int(min(65536, (1 << 30) // (16 * n), max(use_threads * 4, target_seconds * use_threads / one)))
        x = np.ascontiguousarray(np.tile(probe, (frames // probe.shape[0] + 1, 1))[:frames])
        t0 = time.perf_counter()
        O.fft(x, radix, False, impl, threads=use_threads)
        dt = time.perf_counter() - t0
        return (frames * n / dt / 1e6, kind, use_threads,
                f"{frames} frames x {n}-pt fft_radix{radix} fp64 (the reference is fp64-only), g++ -O3 -DNDEBUG, {use_threads} threads, {dt:.2f} s",
                n / one / 1e6)
    ch_total, n_total, m = spec["channels"], spec["samples"], spec["sections"]
    n = min(n_total, 1 << 18)
    ch = min(ch_total, max(threads * 2, 16))
    x = rng.standard_normal((ch, n))
    ftype = np.where(np.arange(ch) % 2 == 0, 1, 2)
    f0 = np.geomspace(1e3, 20e3, ch)
    if kind == "reference":
        t0 = time.perf_counter()
        O.iir_bank_reference(x[:1], ftype[:1], f0[:1], 100e3, sections=m, threads=1)
        one = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.iir_bank_reference(x, ftype, f0, 100e3, sections=m, threads=threads)  # also sizes the sample
        reps = max(1, int(target_seconds / max(time.perf_counter() - t0, 1e-3)))
        t0 = time.perf_counter()
        for _ in range(reps):
            O.iir_bank_reference(x, ftype, f0, 100e3, sections=m, threads=threads)
        dt = time.perf_counter() - t0
        use_threads = min(threads, ch)
    else:
        t0 = time.perf_counter()
        O.iir_bank_port(x[:1], ftype[:1], f0[:1], 100e3, sections=m)
        one = time.perf_counter() - t0
        reps = 1
        t0 = time.perf_counter()
        O.iir_bank_port(x, ftype, f0, 100e3, sections=m)
        dt = time.perf_counter() - t0
        use_threads = 1
    return (reps * ch * n / dt / 1e6, kind, use_threads,
            f"{reps} x {ch} channels x {n} samples casc_2o_iir<{m}> fp64, one object per channel, {use_threads} threads, {dt:.2f} s",
            n / one / 1e6)


def run_reference(args, spec, workload_name):
    """--impl reference: the reference's own CPU implementation (oracle/_ref = the unmodified headers compiled by
    oracle/Makefile) on this box's host cores, same `config`, `metric`, `unit`; every step is a bounded sample of the
    workload sized so that W + K steps end within about 90 s.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    warmup, steps = max(args.warmup, 0), max(args.steps, 1)
    per_step = min(4.0, max(0.4, 90.0 / (warmup + steps)))
    for _ in range(warmup):
        cpu_reference_rate(spec, threads, target_seconds=per_step)
    rates, last = [], None
    t0 = time.perf_counter()
    for _ in range(steps):
        last = cpu_reference_rate(spec, threads, target_seconds=per_step)
        rates.append(last[0])
    wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": wall / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(workload_name, spec),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": last[2], "kind": last[1], "sample": last[3],
                         "single_thread": last[4]},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "details": {"dtype_note": "the reference computes in fp64 whatever the workload's precision; the b200 arm's fp64 "
                                  "numbers are in its `secondary` entries fft4096_f64 / iir16384_f64"},
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------
# (the short FFT workloads first: measured right after the IIR banks or config 5, which run into the power cap for seconds, their first
# steps still see the lowered clocks -- profiles/r02_bench_default_v7.json caught fft8192_f32 at 1.0 ms against 0.81)
SECONDARY = ("fft4096_f64", "fft8192_f32", "fft16384_f32", "fftr2c4096_f32", "fft65536_f32", "iir16384_f32", "iir16384_f32_scan", "iir4096_f32",
             "iirscan_f64", "pipeline_cfg5_f32")
E2E_SECONDARY = ("fft4096_f64", "iir16384_f32", "fftr2c4096_f32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fft4096_f32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="default workload only: skip the other BASELINE configs measured beside it")
    ap.add_argument("--frames", type=int, default=0, help="override frames (fft) for quick runs")
    args = ap.parse_args()
    spec = dict(WORKLOADS[args.workload])
    if args.frames and spec["kind"] == "fft":
        spec["frames"] = args.frames

    if args.impl == "reference":
        run_reference(args, spec, args.workload)
        return

    import torch

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    rank, local, world = dist_setup(args.gpus)
    peak, peak_src = hbm_peak()

    def build(spec):
        if spec["kind"] == "pipeline":
            return PipelineWorkload(spec, local, world)
        return {"fft": FftWorkload, "iir": IirWorkload}[spec["kind"]](spec, local)

    def measure(name, spec, steps):
        """W warm-up steps, then `steps` timed steps between barriers; CUDA events on the launching stream, max over ranks."""
        wl = build(spec)
        for _ in range(warmup):
            wl.step()
        barrier(world)
        sampler = ClockSampler(local)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sampler.start()
        barrier(world)
        evs[0].record()
        for i in range(steps):
            wl.step()
            evs[i + 1].record()
        barrier(world)
        clocks = sampler.stop()
        total_ms = evs[0].elapsed_time(evs[-1])
        per_step = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)])
        total_ms_max = max_over_ranks(total_ms, world)
        check = wl.self_check()
        launches = wl.launches_per_step()
        kernel_ms = float(per_step.mean()) / max(1, launches)
        alg_bytes = wl.samples_per_step * spec["bytes_per_sample"] / max(1, launches)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        res = {
            "value": wl.samples_per_step * steps * world / (total_ms_max * 1e-3) / 1e6, "ms_per_step": total_ms_max / steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get(name, (None, None))[0], "traffic_source": NCU_TRAFFIC.get(name, (None, None))[1],
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": kernel_ms},
            "gpu_launches": int(getattr(wl, "gpu_launches_per_step", wl.launches_per_step)() * steps), "clocks": clocks,
            "self_check": {{"fft": "roundtrip_rel_err", "iir": "nonfinite", "pipeline": "spectrum_vs_torch_fft_rel_err"}[spec["kind"]]: check},
            "step_ms": {"min": float(per_step.min()), "median": float(np.median(per_step)), "max": float(per_step.max())},
            "plan": wl.describe(),
        }
        return wl, res

    def measure_e2e(wl, spec, steps):
        """The same metric through the public host-buffer API: pinned host memory, H2D + D2H inside the timed region."""
        h2d, d2h = wl.e2e_prepare()
        e2e_steps = max(2, min(steps, 6))
        wl.e2e_step()  # warm-up (allocates staging)
        wl.e2e_step()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.e2e_step()
        barrier(world)
        e2e_s = max_over_ranks(time.perf_counter() - t0, world)
        err = wl.e2e_check()
        samples = getattr(wl, "e2e_samples_per_step", wl.samples_per_step)
        wl.e2e_release()
        api = ("simpledsp_b200.FftPlan.half_spectrum(numpy views of pinned host memory) -> sdsp_b200_fft_exec_r2c(PTR_HOST)" if spec.get("half") else
               "simpledsp_b200.FftPlan.__call__(numpy view of pinned host memory) -> sdsp_b200_fft_exec(PTR_HOST)" if spec["kind"] == "fft" else
               "simpledsp_b200.IirBank.process_ptr(pinned host memory) -> sdsp_b200_iir_bank_process(PTR_HOST), staged in chunks of time")
        return {"value": samples * e2e_steps * world / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "samples_per_step": samples,
                "pcie_GBps_each_way": h2d * e2e_steps / e2e_s / 1e9,
                ("roundtrip_rel_err" if spec["kind"] == "fft" else "nonfinite"): err, "api": api}

    wl, res = measure(args.workload, spec, steps)
    value = res["value"]

    # ---- end to end (host buffers through the public API), rank-local, max over ranks
    e2e = None
    if not args.no_e2e and spec["kind"] in ("fft", "iir"):
        e2e = measure_e2e(wl, spec, steps)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if _ORIGINAL_AFFINITY:
            os.sched_setaffinity(0, _ORIGINAL_AFFINITY)  # the CPU baseline uses every host core
        r = cpu_reference_rate(spec, os.cpu_count() or 1)
        cpu = {"value": r[0], "unit": "Msamples/s", "cores": r[2], "kind": r[1], "sample": r[3], "single_thread": r[4]}

    # ---- the rest of the metric ("batched FFT & biquad IIR", every BASELINE config) measured in the same run
    secondary = []
    if args.workload == "fft4096_f32" and not args.no_secondary:
        wl.data = None
        del wl
        torch.cuda.empty_cache()
        for name in SECONDARY:
            sp = dict(WORKLOADS[name])
            sec_steps = 5 if sp["kind"] != "fft" else 20  # (the FFT workloads take about a millisecond per step)
            w2, r2 = measure(name, sp, sec_steps)
            entry = {"workload": name, "metric": "Msamples/s", "value": r2["value"], "unit": "Msamples/s", "ms_per_step": r2["ms_per_step"],
                     "steps": sec_steps, "step_ms": r2["step_ms"], "n_gpus": world, "scaling": "strong" if sp.get("strong") else "weak", "dtype": sp["precision"],
                     "roofline": r2["roofline"], "gpu_launches": r2["gpu_launches"], "self_check": r2["self_check"], "clocks": r2["clocks"],
                     "plan": r2["plan"], "config": config_of(name, sp)}
            if name in E2E_SECONDARY and not args.no_e2e:
                entry["e2e"] = measure_e2e(w2, sp, 4)
            secondary.append(entry)
            for attr in ("data", "signal", "spectra"):
                if hasattr(w2, attr):
                    setattr(w2, attr, None)
            del w2
            torch.cuda.empty_cache()

    if rank == 0:
        out = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if spec.get("strong") else "weak",
            "vs_baseline": None, "dtype": spec["precision"], "data": "synthetic", "config": config_of(args.workload, spec),
            "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "self_check": res["self_check"], "step_ms": res["step_ms"],
            "details": {"plan": res["plan"], "steps_alternate": "forward/reverse" if spec["kind"] == "fft" else "n/a"},
            "secondary": secondary,
        }
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
