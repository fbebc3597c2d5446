"""Import stub: the reference imports trimesh at module level (legged_gym/utils/terrain.py:32) but only
uses it for mesh_type == 'trimesh', which the hot path never selects. Oracle/test tooling only."""


class Trimesh:
    def __init__(self, *a, **k):
        pass
