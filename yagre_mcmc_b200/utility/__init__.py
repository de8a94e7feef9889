from .hierarchy import HierarchyBase, SharedComponent, Hierarchy
