"""Open3D-shaped shim: ``import <pkg>.o3d_compat as o3d`` lets verfication.py / icp.py run
unchanged on the CUDA library.  Only the subset those scripts touch is provided:

  o3d.geometry.PointCloud            verfication.py:81-89,97-99; icp.py:83-86,110-115
  o3d.utility.Vector3dVector         verfication.py:82,87,89; icp.py:84,86
  o3d.io.read_point_cloud            icp.py:112            (ASCII / binary PLY vertices)
  o3d.pipelines.registration.*       icp.py:97-103
  o3d.visualization.draw_geometries  verfication.py:27, icp.py:14,23  (no-op)
"""
from . import geometry, io, pipelines, utility, visualization  # noqa: F401

__all__ = ["geometry", "io", "pipelines", "utility", "visualization"]
