from .evaluation import EvaluationStatus
from .interface import SolverInterface
from .forwardModel import ForwardModel
from .solvers import LinearModelSolver, LotkaVolterraRK4Solver, LotkaVolterraParameter
