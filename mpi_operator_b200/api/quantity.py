"""Kubernetes resource.Quantity arithmetic, enough for PodGroup minResources
(reference: pkg/controller/podgroup.go:420-443 addResources)."""
from __future__ import annotations

from decimal import Decimal
from typing import Dict, Union

_BIN = {"Ki": 2 ** 10, "Mi": 2 ** 20, "Gi": 2 ** 30, "Ti": 2 ** 40, "Pi": 2 ** 50, "Ei": 2 ** 60}
_DEC = {"n": Decimal("1e-9"), "u": Decimal("1e-6"), "m": Decimal("1e-3"), "": Decimal(1), "k": Decimal("1e3"),
        "M": Decimal("1e6"), "G": Decimal("1e9"), "T": Decimal("1e12"), "P": Decimal("1e15"), "E": Decimal("1e18")}


class Quantity:
    """Value + the suffix family it was written in (so '2Gi' + '1Gi' prints '3Gi')."""

    __slots__ = ("value", "fmt")

    def __init__(self, value: Decimal, fmt: str = "dec"):
        self.value, self.fmt = Decimal(value), fmt

    @classmethod
    def parse(cls, s: Union[str, int, float, "Quantity"]) -> "Quantity":
        if isinstance(s, Quantity):
            return Quantity(s.value, s.fmt)
        if isinstance(s, (int, float)):
            return cls(Decimal(str(s)), "dec")
        s = s.strip()
        for suf, mul in _BIN.items():
            if s.endswith(suf):
                return cls(Decimal(s[:-2]) * mul, "bin")
        if s and s[-1] in "numkMGTPE" and not s[-1].isdigit():
            return cls(Decimal(s[:-1]) * _DEC[s[-1]], "dec")
        return cls(Decimal(s), "dec")

    def __add__(self, o: "Quantity") -> "Quantity":
        return Quantity(self.value + o.value, self.fmt)

    def __mul__(self, k: int) -> "Quantity":
        return Quantity(self.value * k, self.fmt)

    def __eq__(self, o) -> bool:
        return isinstance(o, Quantity) and self.value == o.value

    def __lt__(self, o: "Quantity") -> bool:
        return self.value < o.value

    def __le__(self, o: "Quantity") -> bool:
        return self.value <= o.value

    def __int__(self) -> int:
        return int(self.value)

    def __str__(self) -> str:
        v = self.value
        if self.fmt == "bin":
            for suf in ("Ei", "Pi", "Ti", "Gi", "Mi", "Ki"):
                m = _BIN[suf]
                if v >= m and v % m == 0:
                    return f"{int(v // m)}{suf}"
            return str(int(v)) if v == v.to_integral() else str(v)
        if v == v.to_integral():
            iv = int(v)
            for suf in ("E", "P", "T", "G", "M", "k"):
                m = int(_DEC[suf])
                if iv >= m and iv % m == 0:
                    return f"{iv // m}{suf}"
            return str(iv)
        milli = v * 1000
        if milli == milli.to_integral():
            return f"{int(milli)}m"
        return str(v)

    def __repr__(self) -> str:
        return f"Quantity({str(self)!r})"


ResourceList = Dict[str, str]


def add_to(total: Dict[str, Quantity], name: str, q: Quantity) -> None:
    total[name] = total[name] + q if name in total else Quantity(q.value, q.fmt)


def to_strings(rl: Dict[str, Quantity]) -> ResourceList:
    return {k: str(v) for k, v in rl.items()}
