from .diagnostics import DummyDiagnostics, AcceptanceRateDiagnostics, FullDiagnostics
from .target import UnnormalisedPosterior
