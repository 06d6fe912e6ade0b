"""Small driver that launches K4 (quad demod rows), K5 (convert_iq) and K6 (post-demod chains) once with realistic sizes
(for ncu captures)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from radiocapture_rf_b200 import _lib
from radiocapture_rf_b200.engine import Engine, PostDemod

e = Engine(0)
rng = np.random.default_rng(0)
rows, n = 256, 25000
x = (np.exp(1j * np.cumsum(rng.uniform(-0.3, 0.3, (rows, n)), axis=1)) * 0.5).astype(np.complex64)
fm = e.quad_demod(x, 5.0) if hasattr(e, "quad_demod") else None
raw = rng.integers(0, 255, 2 << 24, dtype=np.uint8)
if hasattr(e, "convert_iq"):
    e.convert_iq(raw, "u8")
pd = PostDemod.p25_c4fm(e, rows)
pd.process(x)
pa = PostDemod.analog_fm(e, 64)
pa.process(x[:64])
e.close()
print("k456 driver done")
