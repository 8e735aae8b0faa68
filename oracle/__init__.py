"""CPU oracle for the audio-diffusion hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of this path lives in `diffusers==0.24.0` and
`librosa==0.10.2.post1` (requirements-lock.txt:25,57 of the reference), neither
of which is vendored under /root/reference nor installable here (no network).
The reference itself carries no tests, fixtures or golden vectors for the path
(SURVEY.md §4).  Every function in this package is therefore a restatement of
the published algorithm of the pinned third-party version, anchored on the
reference's own call sites (cited per function).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; the product (`audio_diffusion_b200`) never does.
"""
