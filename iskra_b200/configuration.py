"""problem/configuration.jl:5-15,141 -- the untyped Config bag the runner fills (src/iskra.jl:43-52)."""


class Config:
    def __init__(self):
        self.solver = None
        self.pusher = None
        self.tracker = None
        self.interactions = []
        self.species = []
        self.sources = []
        self.circuit = None
        self.grid = None
        self.cells = None
