"""B200-native registration-and-verification hot path (Chamfer / ADD-S candidate scoring
and point-to-point ICP) behind the reference's call surface.

Public surface
  api        batched device API: transform_points, nearest_neighbors, chamfer_distance,
             adds, verify_poses, evaluate_registration, icp, multistart_icp, icp_refine_pose
  helpers    the reference scripts' helper names (calculate_relative_pose, ADD, ADDS, ...)
  o3d_compat Open3D-shaped shim (``import ...o3d_compat as o3d``)
  compat     sklearn-shaped KDTree shim
  dist       candidate / ICP sharding over the GPUs of one box (torch.distributed + NCCL)
  synth      seeded synthetic clouds and candidate poses (tests, bench)

All arithmetic runs in csrc/libisr.so (hand-written sm_100a CUDA behind the C ABI of
include/isr.h).  There is no CPU fallback.
"""
from . import _lib, synth  # noqa: F401
from .api import (IcpProblem, IcpResult, MultiStartResult, NNResult, SoaCloud,  # noqa: F401
                  VerifyResult, adds, chamfer_distance, evaluate_registration, icp,
                  centroid_of, measure_fp32_peak, multistart_icp, nearest_neighbors, pack_soa,
                  prepare_cloud, spatial_order, set_nn_pruning, get_nn_pruning, radius_neighbor_count,
                  point_cloud_distance, pose_from_Rt, icp_refine_pose, refine_pose, score_pnp_hypotheses, adds_rigid, vote, pnp_ransac, p3p_solve, estimate_normals, transform_points,
                  verify_poses)
from .helpers import (ADD, ADDS, calculate_relative_pose, choose_image, choose_image_from_poses,  # noqa: F401
                      compute_rel_poses, draw_registration_result, relative_pose_table,
                      pnp, select_pnp_hypothesis, vp)

__version__ = "0.1.0"
