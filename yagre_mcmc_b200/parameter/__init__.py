from .interface import ParameterInterface
from .vector import ParameterVector
from .scalar import ScalarParameter
