"""ObjTracker/utils/constants.py values the hot path reads."""
REND_SIZE = 256  # constants.py:2  size of target masks for the silhouette loss
BBOX_EXPANSION_FACTOR = 0.3  # constants.py:3
