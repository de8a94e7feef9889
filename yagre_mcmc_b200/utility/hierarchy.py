"""Level containers (reference: yagremcmc/utility/hierarchy.py:5-63). level(-1) is the finest."""
from abc import ABC, abstractmethod


class HierarchyBase(ABC):

    def __init__(self, nLevels):
        if nLevels < 1:
            raise ValueError(f"Trying to set up hierarchy with {nLevels} levels.")
        self._nLevels = nLevels

    @property
    def size(self):
        return self._nLevels

    def validate_level_index(self, idx):
        if idx < -1 or self._nLevels <= idx:
            raise ValueError(f"invalid level index. Trying to access level {idx} in a hierarchy of "
                             f"{self._nLevels} levels.")

    @abstractmethod
    def level(self, lvlIdx):
        ...


class SharedComponent(HierarchyBase):
    """One object shared by every level (prior, noise, data)."""

    def __init__(self, sharedComponent, nLevels):
        super().__init__(nLevels)
        self._sharedComponent = sharedComponent

    def level(self, lvlIdx):
        self.validate_level_index(lvlIdx)
        return self._sharedComponent


class Hierarchy(HierarchyBase):

    def __init__(self, hierarchy):
        super().__init__(len(hierarchy))
        self._hierarchy = list(hierarchy)

    def level(self, lvlIdx):
        self.validate_level_index(lvlIdx)
        return self._hierarchy[lvlIdx]          # -1 indexes the finest level
