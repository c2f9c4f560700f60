"""Back-compat alias: the seeded input generator lives in /synth_inputs.py (it is not part of the
checker; bench.py and smoke() import it from there).  Tests may keep `from oracle import synth`."""
from synth_inputs import *  # noqa: F401,F403
from synth_inputs import _rodrigues, _skew  # noqa: F401
