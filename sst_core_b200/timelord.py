"""Time/frequency strings -> integer core cycles.

Restates the subset of TimeLord / UnitAlgebra the hot path needs
(reference: src/sst/core/timeLord.cc:139-175 getFactorForTime, src/sst/core/unitAlgebra.cc):
a value with an SI prefix and a unit of seconds ("1ns", "10 us", "2.5ms", "10000s") or Hertz ("1GHz",
"500MHz") is turned into a number of core cycles of the time base (default "1ps", config.cc:490):
    seconds:  value / timebase            Hertz:  (1/timebase) / value
rounded to nearest (UnitAlgebra::getRoundedValue); a non-zero result below one cycle is an error
(std::underflow_error in the reference).
"""
from __future__ import annotations

import re
from fractions import Fraction

_SI = {
    "a": Fraction(1, 10**18), "f": Fraction(1, 10**15), "p": Fraction(1, 10**12), "n": Fraction(1, 10**9),
    "u": Fraction(1, 10**6), "m": Fraction(1, 10**3), "": Fraction(1), "k": Fraction(10**3),
    "K": Fraction(10**3), "M": Fraction(10**6), "G": Fraction(10**9), "T": Fraction(10**12),
    "P": Fraction(10**15), "E": Fraction(10**18),
}
_RE = re.compile(r"^\s*([0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)\s*([afpnumkKMGTPE]?)\s*(s|Hz|hz)\s*$")


class TimeError(ValueError):
    pass


def parse(text) -> tuple[Fraction, str]:
    """Return (value in base units, 's' | 'Hz')."""
    m = _RE.match(str(text))
    if not m:
        raise TimeError(f"Error: TimeConverter creation requires a time unit (s or Hz), '{text}' was passed")
    val = Fraction(m.group(1)) * _SI[m.group(2)]
    unit = "Hz" if m.group(3).lower() == "hz" else "s"
    return val, unit


def _round(fr: Fraction) -> int:
    return int(fr + Fraction(1, 2))  # round half up, values are non-negative


class TimeLord:
    def __init__(self, timebase: str = "1ps"):
        tb, unit = parse(timebase)
        if unit != "s" or tb <= 0:
            raise TimeError("timebase must be a positive time")
        self.timebase = tb
        self.timebase_string = timebase

    def cycles(self, text) -> int:
        """timeLord.cc:139-175."""
        val, unit = parse(text)
        if unit == "s":
            factor = val / self.timebase
        else:
            if val == 0:
                raise TimeError("zero frequency")
            factor = (1 / self.timebase) / val
        if factor > (1 << 64) - 1:
            raise OverflowError(f"time {text} too large for timebase {self.timebase_string}")
        if factor != 0 and factor < 1:
            raise TimeError(f"time {text} has too small a period for timebase {self.timebase_string}")
        return _round(factor)

    def best_si(self, cycles: int) -> str:
        """'Simulation is complete, simulated time: 10 us' formatting (UnitAlgebra::toStringBestSI)."""
        t = Fraction(cycles) * self.timebase
        if t == 0:
            return "0 s"
        for pre in ["", "m", "u", "n", "p", "f", "a"]:
            scale = _SI[pre]
            if t >= scale:
                v = t / scale
                txt = f"{float(v):.6g}"
                return f"{txt} {pre}s"
        return f"{float(t / _SI['a']):.6g} as"
