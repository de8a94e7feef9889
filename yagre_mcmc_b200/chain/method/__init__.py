from .mrw import MRWBuilder, MetropolisedRandomWalk
from .mlda import MLDABuilder, MLDA
from .am import AMBuilder, AdaptiveMetropolis
from .pcn import PCNBuilder, PreconditionedCrankNicolson
from .aem import AEMBuilder, AdaptiveErrorModel
