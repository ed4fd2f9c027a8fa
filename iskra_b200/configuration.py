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


def create_fluid_species(name, mu, q, m, mx, my, T=300.0):
    """create_fluid_species  problem/configuration.jl:104-108"""
    import numpy as np

    from . import particle_in_cell as PIC
    return PIC.FluidSpecies(name, mu, q, m, np.zeros((mx, my)), T)


def create_electrode(nodes, config_or_solver, grid=None, fixed=False, sigma=0.0, phi=0.0):
    """create_electrode  problem/configuration.jl:22-72.  Two call forms like the reference:
    create_electrode(nodes, config; ...) also registers the electrode with config.tracker (creating the
    tracker on first use, :31-33); create_electrode(nodes, solver, grid; ...) only edits the solver."""
    import numpy as np

    from . import finite_difference_method as FDM
    from . import particle_in_cell as PIC
    config = None
    if grid is None:
        config = config_or_solver
        if config.grid is None:
            print("No grid defined. EXIT")                       # :24-26
        if config.solver is None:
            print("No field solver defined. EXIT")               # :27-29
        if config.tracker is None:
            config.tracker = PIC.create_surface_tracker(config.grid)   # :31-33
        ps, grid = config.solver, config.grid
    else:
        ps = config_or_solver
    nodes = np.asarray(nodes, dtype=bool)
    nx, ny = grid.n
    dx, dy = grid.dh
    area = 0.0                                                   # calculate_area :45-53 (dz = 1)
    for j in range(ny):
        for i in range(nx):
            if nodes[i, j]:
                area += dx * 1.0 if (i + 1 < nx and nodes[i + 1, j]) else 0
                area += dy * 1.0 if (j + 1 < ny and nodes[i, j + 1]) else 0
    k = int(np.flatnonzero(nodes.reshape(-1, order="F"))[0])     # find_reference_node :54-57
    i, j = k % nx + 1, k // nx + 1
    if fixed:
        FDM.apply_dirichlet(ps, nodes, phi)                      # :62
        el = PIC.FixedPotentialElectrode(FDM.DirichletRhs(phi), 0.0, area)
    else:
        dof = FDM.add_new_dof(ps, "sigma")                       # :66
        FDM.apply_neumann(ps, nodes, dof)                        # :67
        s0 = FDM.get_rhs(ps, "sigma", dof)                       # :68
        s0.value = sigma
        p0 = FDM.get_solution(ps, "phi", i, j)                   # :69
        el = PIC.FloatingPotentialElectrode(p0, s0, 0.0, area)   # :70 (argument order as in the reference, S1)
        el._dof = dof
    if config is not None:
        PIC.track_surface_(config.tracker, nodes, el)            # :37
    return el
