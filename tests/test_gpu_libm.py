"""The device build of csrc/elmk_libm.h against the host's libm: elmk_math_eval runs the library's own m_exp / m_log /
... on the GPU at given arguments; the checker library (oracle/_ref or the port) answers the same call with libm.
Bit-for-bit equality is required for every function (exp, log, log10, pow, atan, cos, acos, tanh, erf) and for division."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 2_000_000


def cases(rng):
    u = rng.uniform
    lg = lambda e0, e1, n: np.ldexp(u(1.0, 2.0, n), rng.integers(e0, e1, n))
    sg = lambda a: a * rng.choice([-1.0, 1.0], a.size)
    yield "exp", u(-40, 40, N), None
    yield "exp", u(-745, 710, N), None
    yield "exp", sg(lg(-60, 10, N)), None
    yield "log", u(1e-300, 4, N), None
    yield "log", u(0.9, 1.1, N), None
    yield "log", lg(-1022, 1023, N), None
    yield "log10", u(1e-300, 2000, N), None
    yield "log10", lg(-300, 300, N), None
    yield "atan", u(-16, 16, N), None
    yield "atan", sg(lg(-40, 60, N)), None
    yield "cos", u(-np.pi, np.pi, N), None
    yield "cos", u(-1000, 1000, N), None
    yield "tanh", u(-3, 3, N), None
    yield "tanh", sg(lg(-60, 6, N)), None
    yield "erf", u(-7, 7, N), None
    yield "erf", sg(lg(-60, 4, N)), None
    yield "acos", u(-1, 1, N), None
    yield "acos", 1.0 - lg(-52, -3, N), None
    yield "acos", sg(lg(-60, 0, N)), None
    yield "pow", u(1e-6, 2, N), u(-3, 3, N)
    yield "pow", u(1e-9, 1e3, N), u(0, 1, N)
    yield "pow", lg(-200, 200, N), u(-8, 8, N)
    yield "pow", u(1e-8, 400, N), rng.choice([3.0, 4.0, 0.333, 0.45, 0.25, 1.5, 0.666666666666, -0.5, 0.5, -0.333], N)
    yield "pow", np.full(N, 2.0), u(-12, 12, N)
    yield "div", sg(lg(-300, 300, N)), sg(lg(-300, 300, N))
    yield "div", np.where(rng.uniform(size=N) < 0.3, 0.0, u(-5, 5, N)), sg(lg(-30, 30, N))


def test_device_transcendentals_equal_host_libm(cuda_lib, checker):
    a = cuda_lib.columns(128)
    b = checker.columns(128)
    rng = np.random.default_rng(20241018)
    for fn, x, y in cases(rng):
        got = a.math_eval(fn, x, y)
        ref = b.math_eval(fn, x, y)
        same = (got.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(got) & np.isnan(ref))
        bad = np.nonzero(~same)[0]
        assert bad.size == 0, (f"{fn}: {bad.size} of {x.size} results differ from libm; first x={x[bad[0]]!r}"
                               + (f" y={y[bad[0]]!r}" if y is not None else "")
                               + f" device {got[bad[0]].hex()} libm {ref[bad[0]].hex()}")
