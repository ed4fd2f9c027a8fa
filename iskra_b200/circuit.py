"""Host mirror of Circuit/src/Circuit.jl -- a scalar recurrence per step; it stays on the host (the
reference's circuit is one CircuitRLC advanced once per time step, ParticleInCell.jl:116)."""


class CircuitDevice:
    """abstract type CircuitDevice  Circuit.jl:12"""

    def voltage(self):
        return 0.0


class ShortedConnection(CircuitDevice):
    """Circuit.jl:25, voltage :52"""


class CircuitRLC:
    """mutable struct CircuitRLC  Circuit.jl:14-23; CircuitRLC(i0, q0, t0) :27-28"""

    def __init__(self, i0=0.0, q0=0.0, t0=0.0):
        self.R = self.L = self.C = 0.0
        self.i, self.q, self.t = float(i0), float(q0), float(t0)
        self.V = lambda t: 0.0
        self.ext = ShortedConnection()
        self.probes = {}


class Resistor:
    def __init__(self, name, val):
        self.name, self.val = name, float(val)

    def assign(self, cir):
        cir.R = self.val                                  # :53


class Inductor:
    def __init__(self, name, val):
        self.name, self.val = name, float(val)

    def assign(self, cir):
        cir.L = self.val                                  # :54


class Capacitor:
    def __init__(self, name, val):
        self.name, self.val = name, float(val)

    def assign(self, cir):
        cir.C = self.val                                  # :55


class VoltageSource:
    def __init__(self, name, val):
        self.name, self.val = name, val

    def assign(self, cir):
        cir.V = self.val                                  # :56


class ExternalDevice:
    def __init__(self, name, val):
        self.name, self.val = name, val

    def assign(self, cir):
        cir.ext = self.val                                # :57


def netlist(entries):
    """@netlist begin NAME, node+, node-, value ... end  Circuit.jl:73-115: the element kind follows
    from the first letters of the name (V, R, L, C, EXT)."""
    out = []
    for name, _np, _nm, value in entries:
        if name.startswith("V"):
            out.append(VoltageSource(name, value))
        elif name.startswith("R"):
            out.append(Resistor(name, value))
        elif name.startswith("L"):
            out.append(Inductor(name, value))
        elif name.startswith("C"):
            out.append(Capacitor(name, value))
        elif name.startswith("EXT"):
            out.append(ExternalDevice(name, value))
        else:
            raise ValueError("unknown element")           # :114
    return out


def rlc(elements):
    """rlc(elements)  Circuit.jl:58-66"""
    cir = CircuitRLC(0, 0, 0)
    for e in elements:
        e.assign(cir)
    return cir


def damping_factor(R, L, C):
    return (R / (2 * L)) * (L * C) ** 0.5                 # :67


def resonant_frequency(L, C):
    import math
    return 1 / math.sqrt(L * C) / (2 * math.pi)           # :69


def advance_circuit_(cir, V, dt):
    """advance_circuit!(cir, V, dt)  Circuit.jl:117-136"""
    t, v = cir.t, cir.V
    i, q = cir.i, cir.q
    R, L, C = cir.R, cir.L, cir.C
    vext = cir.ext.voltage()
    cir.i = (L / dt - R / 2) * i + vext - v(t)
    if C > 0.0:
        cir.i -= q / C
        cir.i /= (L / dt + R / 2)
        cir.q = q + dt * i
    else:
        cir.i /= (L / dt + R / 2)
    cir.t += dt
    cir.probes = {"Q1": cir.q, "I1": cir.i, "V1": v(t), "Vext": vext}   # @probe :132-135
    from . import diagnostics as DG
    for key, units in (("Q1", "C"), ("I1", "A"), ("V1", "V"), ("Vext", "V")):
        if not isinstance(DG.records.get(key), DG.ProbeRecord) or DG.records[key].owner is not cir:
            DG.register_probe(key, (lambda k=key: cir.probes[k]), units)
            DG.records[key].owner = cir
