"""Fixture simulators written against the public Madrona API.

The reference's example simulators (Escape Room, Hide&Seek, Cartpole,
Overcooked) live in other repositories that are not available here
(SURVEY.md F7), so these self-authored sims define the BASELINE.json workloads
by construction.  Each sim's C++ sources are compiled twice from the same
files: by g++ against the reference headers + CPU backend (oracle/) and by
NVRTC against madrona_b200/device (the B200 engine).
"""
from .fixtures import SIMS, SimDesc, make_executor, pack_world_inits  # noqa: F401
