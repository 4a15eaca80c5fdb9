from .run import (CashStore, EmptyStore, Simulation, init_particles, initialize_simulation,  # noqa: F401
                  reset_simulation, run)
