from .covariance import DiagonalCovarianceMatrix, IIDCovarianceMatrix, DenseCovarianceMatrix
from .gaussian import Gaussian, GaussianDensity
from .noise import CentredGaussianNoise
from .data import Data
from .likelihood import AdditiveGaussianNoiseLikelihood
from .bayesModel import BayesianRegressionModel
from .modelHierarchy import BayesianRegressionModelHierarchy
from .estimation import WelfordAccumulator
from .targets import GaussianTargetDensity, GaussianTargetDensity1d, GaussianTargetDensity2d
