from .covariance import DiagonalCovarianceMatrix, IIDCovarianceMatrix, DenseCovarianceMatrix
from .gaussian import Gaussian, GaussianDensity
from .noise import CentredGaussianNoise, AEMNoise
from .data import Data
from .likelihood import AdditiveGaussianNoiseLikelihood, AEMLikelihood
from .bayesModel import BayesianRegressionModel
from .modelHierarchy import BayesianRegressionModelHierarchy
from .estimation import WelfordAccumulator
from .targets import GaussianTargetDensity, GaussianTargetDensity1d, GaussianTargetDensity2d
