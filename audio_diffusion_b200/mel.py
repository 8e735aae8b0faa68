"""`Mel` — same constructor, attributes and methods as audiodiffusion/mel.py:44-168
(`load_audio`, `get_number_of_slices`, `get_audio_slice`, `get_sample_rate`, `set_resolution`,
`audio_slice_to_image`, `image_to_audio`), with the librosa arithmetic replaced by the batched CUDA kernels of
libb200ad.so (`b200ad_mel_encode` / `b200ad_mel_decode`).  Batched variants (`audio_slices_to_images`,
`images_to_audio`) are what the pipeline and the dataset builder use.

Host-side numpy is used only for data-independent constants (the Slaney mel filterbank of librosa.filters.mel
and its pseudo-inverse) and for file decoding in `load_audio`.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Callable, List, Optional, Union

import numpy as np
import torch
from PIL import Image

from . import _lib
from ._lib import MelConfigC


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=float)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, dtype=np.float32) -> np.ndarray:
    """The constant librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels) produces (Slaney scale and norm)."""
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - freqs[None, :]
    lower = -ramps[:-2] / width[:-1, None]
    upper = ramps[2:] / width[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper)).astype(dtype)
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class Mel:
    config_name = "mel_config.json"

    def __init__(self, x_res: int = 256, y_res: int = 256, sample_rate: int = 22050, n_fft: int = 2048,
                 hop_length: int = 512, top_db: int = 80, n_iter: int = 32):
        self.hop_length = hop_length
        self.sr = sample_rate
        self.n_fft = n_fft
        self.top_db = top_db
        self.n_iter = n_iter
        self.set_resolution(x_res, y_res)
        self.audio = None
        self.config = _Cfg(x_res=x_res, y_res=y_res, sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length,
                           top_db=top_db, n_iter=n_iter, _class_name="Mel")
        self.phase_seed = 0
        self._dev_cache = {}

    # ------------------------------------------------------------------ reference API (mel.py:80-133)
    def set_resolution(self, x_res: int, y_res: int):
        self.x_res = x_res
        self.y_res = y_res
        self.n_mels = self.y_res
        self.slice_size = self.x_res * self.hop_length - 1
        self._dev_cache = {}

    def load_audio(self, audio_file: str = None, raw_audio: np.ndarray = None):
        if audio_file is not None:
            self.audio = _decode_audio_file(audio_file, self.sr)
        else:
            self.audio = raw_audio
        if len(self.audio) < self.x_res * self.hop_length:  # pad with silence (float64, as np.zeros in the reference)
            self.audio = np.concatenate([self.audio, np.zeros((self.x_res * self.hop_length - len(self.audio),))])

    def get_number_of_slices(self) -> int:
        return len(self.audio) // self.slice_size

    def get_audio_slice(self, slice: int = 0) -> np.ndarray:
        return self.audio[self.slice_size * slice: self.slice_size * (slice + 1)]

    def get_sample_rate(self) -> int:
        return self.sr

    # ------------------------------------------------------------------ device plumbing
    def _cfg_c(self) -> MelConfigC:
        return MelConfigC(self.x_res, self.y_res, self.sr, self.n_fft, self.hop_length, self.top_db, self.n_iter)

    def _device(self, device=None) -> torch.device:
        _lib.require_cuda()
        return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    def _constants(self, dev):
        key = str(dev)
        if key not in self._dev_cache:
            basis32 = slaney_mel_basis(self.sr, self.n_fft, self.n_mels, np.float32)
            basis64 = slaney_mel_basis(self.sr, self.n_fft, self.n_mels, np.float64)  # mel_to_stft: dtype=M.dtype
            pinv = np.linalg.pinv(basis64)                                             # (F, n_mels) fp64
            self._dev_cache[key] = (torch.from_numpy(np.ascontiguousarray(basis32.T)).to(dev),
                                    torch.from_numpy(np.ascontiguousarray(pinv)).to(dev))
        return self._dev_cache[key]

    def _scratch(self, n: int, dev) -> torch.Tensor:
        cfg = self._cfg_c()
        need = _lib.lib().b200ad_mel_scratch_bytes(C.byref(cfg), n)
        if need == 0:
            raise _lib.B200ADError(_lib.lib().b200ad_last_error().decode())
        return torch.empty(need, dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------ batched codec (GPU)
    def audio_slices_to_images(self, slices: Union[np.ndarray, torch.Tensor], device=None, ref=None) -> torch.Tensor:
        """(n, slice_size) audio -> (n, y_res, x_res) uint8 on the device (mel.py:145-149, batched).
        `ref` = None / np.max: `librosa.power_to_db(S, ref=np.max)`, the reference default; a number, or a callable evaluated
        on every slice's mel power spectrogram S (what power_to_db does with a callable), otherwise."""
        if ref is not None and ref is not np.max:
            return self._encode_with_ref(slices, device, ref)
        dev = self._device(device)
        a = torch.as_tensor(np.ascontiguousarray(slices) if isinstance(slices, np.ndarray) else slices)
        a = a.to(device=dev, dtype=torch.float32).contiguous()
        if a.ndim != 2 or a.shape[1] != self.slice_size:
            raise ValueError(f"expected (n, {self.slice_size}) audio slices, got {tuple(a.shape)}")
        n = a.shape[0]
        basis_t, _ = self._constants(dev)
        out = torch.empty((n, self.y_res, self.x_res), dtype=torch.uint8, device=dev)
        scratch = self._scratch(n, dev)
        cfg = self._cfg_c()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200ad_mel_encode(C.byref(cfg), basis_t.data_ptr(), a.data_ptr(), out.data_ptr(), n,
                                                    scratch.data_ptr(), scratch.numel(), _lib.stream_ptr()))
        return out

    def _encode_with_ref(self, slices, device, ref) -> torch.Tensor:
        dev = self._device(device)
        a = torch.as_tensor(np.ascontiguousarray(slices) if isinstance(slices, np.ndarray) else slices)
        a = a.to(device=dev, dtype=torch.float32).contiguous()
        if a.ndim != 2 or a.shape[1] != self.slice_size:
            raise ValueError(f"expected (n, {self.slice_size}) audio slices, got {tuple(a.shape)}")
        n = a.shape[0]
        basis_t, _ = self._constants(dev)
        out = torch.empty((n, self.y_res, self.x_res), dtype=torch.uint8, device=dev)
        scratch = self._scratch(n, dev)
        cfg = self._cfg_c()
        L = _lib.lib()
        with torch.cuda.device(dev):
            if callable(ref):      # ref(S) per slice, on the host like librosa (S: float32 (y_res, x_res))
                power = torch.empty((n, self.y_res, self.x_res), dtype=torch.float32, device=dev)
                _lib.check(L.b200ad_mel_encode_ref(C.byref(cfg), basis_t.data_ptr(), a.data_ptr(), None, n, None,
                                                   power.data_ptr(), scratch.data_ptr(), scratch.numel(), _lib.stream_ptr()))
                refs = torch.tensor([float(ref(p)) for p in power.cpu().numpy()], dtype=torch.float32, device=dev)
            else:
                refs = torch.full((n,), float(ref), dtype=torch.float32, device=dev)
            _lib.check(L.b200ad_mel_encode_ref(C.byref(cfg), basis_t.data_ptr(), a.data_ptr(), out.data_ptr(), n,
                                               refs.data_ptr(), None, scratch.data_ptr(), scratch.numel(), _lib.stream_ptr()))
        return out

    def images_to_audio(self, images: Union[np.ndarray, torch.Tensor], device=None) -> np.ndarray:
        """(n, y_res, x_res) uint8 -> (n, (x_res-1)*hop_length) float32 audio (mel.py:162-167, batched)."""
        dev = self._device(images.device if torch.is_tensor(images) and images.is_cuda else device)
        b = torch.as_tensor(images).to(device=dev, dtype=torch.uint8).contiguous()
        if b.ndim != 3 or b.shape[1] != self.y_res or b.shape[2] != self.x_res:
            raise ValueError(f"expected (n, {self.y_res}, {self.x_res}) uint8 images, got {tuple(b.shape)}")
        n = b.shape[0]
        _, pinv = self._constants(dev)
        out = torch.empty((n, (self.x_res - 1) * self.hop_length), dtype=torch.float32, device=dev)
        scratch = self._scratch(n, dev)
        cfg = self._cfg_c()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200ad_mel_decode(C.byref(cfg), pinv.data_ptr(), b.data_ptr(), out.data_ptr(), n,
                                                    C.c_uint64(self.phase_seed), scratch.data_ptr(), scratch.numel(),
                                                    _lib.stream_ptr()))
        self.phase_seed += 1  # the reference draws a fresh (unseeded) random phase on every call
        return out.cpu().numpy()

    # ------------------------------------------------------------------ reference API (mel.py:135-168)
    def audio_slice_to_image(self, slice: int, ref: Union[float, Callable] = np.max) -> Image.Image:
        y = np.asarray(self.get_audio_slice(slice))
        if len(y) != self.slice_size:
            raise ValueError("slice out of range")
        img = self.audio_slices_to_images(y[None, :], ref=ref)[0].cpu().numpy()
        return Image.fromarray(img)

    def image_to_audio(self, image: Image.Image) -> np.ndarray:
        bytedata = np.frombuffer(image.tobytes(), dtype="uint8").reshape((image.height, image.width))
        return self.images_to_audio(bytedata[None].copy())[0]

    # ------------------------------------------------------------------ persistence (mel/mel_config.json)
    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, self.config_name), "w") as f:
            json.dump(dict(self.config), f, indent=2)

    save_config = save_pretrained

    @classmethod
    def from_pretrained(cls, path: str):
        with open(os.path.join(path, cls.config_name)) as f:
            cfg = json.load(f)
        keys = ("x_res", "y_res", "sample_rate", "n_fft", "hop_length", "top_db", "n_iter")
        return cls(**{k: cfg[k] for k in keys if k in cfg})


def _decode_audio_file(path: str, sr: int) -> np.ndarray:
    """Mono float32 audio at `sr`, the contract of `librosa.load(path, mono=True, sr=sr)` (mel.py:100).
    File decoding is outside the hot path; WAV is read with the standard library and resampled with
    scipy.signal.resample_poly."""
    import wave

    if not path.lower().endswith(".wav"):
        raise NotImplementedError("b200 Mel.load_audio decodes WAV files only; pass raw_audio for other formats")
    with wave.open(path, "rb") as w:
        nch, width, rate, nfr = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(nfr)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError("unsupported WAV sample width")
    a = a.reshape(-1, nch).mean(axis=1)
    if rate != sr:
        from math import gcd

        from scipy.signal import resample_poly
        g = gcd(int(rate), int(sr))
        a = resample_poly(a, sr // g, rate // g).astype(np.float32)
    return a
