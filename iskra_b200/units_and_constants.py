"""problem/units_and_constants.jl:3-41 -- Unitful 1.6.0 (CODATA 2018) factors as plain floats."""
import math

K = 1.0
kB = 1.380649e-23
C = 1.0
V = 1.0
m = 1.0
cm = 0.01
u = 1.66053906660e-27
s = 1.0
ms = 1e-3
ns = 1e-9
ps = 1e-12
us = 1e-6
eV = 1.0
kg = 1.0
kHz = 1e3
MHz = 1e6
eps0 = 8.8541878128e-12
mu0 = 1.25663706212e-6
c0 = math.sqrt((1.0 / eps0) * (1.0 / mu0))
qe = 1.6021766208e-19     # units_and_constants.jl:39
me = 9.1093837015e-31     # :40
mp = 1.6726218982e-27     # :41
