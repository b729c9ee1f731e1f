"""dynhor_b200 -- B200-native (sm_100a) joint pose optimisation and DINO template matching for
EAST-J/Dynhor's ObjTracker, behind the reference's own Python call signatures.

    from dynhor_b200.jointopt import joint_optimize, Joint_Optimizer      # ObjTracker/jointopt.py
    from dynhor_b200.losses import Losses, batch_mask_iou                 # ObjTracker/utils/losses.py
    from dynhor_b200.geometry import rot6d_to_matrix, matrix_to_rot6d     # ObjTracker/utils/geometry.py
    import dynhor_b200.neural_renderer as nr                              # the third-party renderer's call surface
    from dynhor_b200.dino_match import dino_cos_topk                      # pose_initializtion.py:295-311

Importing the package does not need a GPU; every compute entry point raises without the CUDA library/device.
"""
__version__ = "0.1.0"
