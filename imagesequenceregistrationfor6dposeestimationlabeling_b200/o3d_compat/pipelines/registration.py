"""o3d.pipelines.registration subset (icp.py:96-103), backed by K1/K2/K3."""
from __future__ import annotations

import numpy as np

from ... import api


class TransformationEstimationPointToPoint:
    def __init__(self, with_scaling: bool = False):
        if with_scaling:
            raise NotImplementedError("with_scaling=True is not used by the reference")
        self.with_scaling = False


class ICPConvergenceCriteria:
    """Defaults as upstream: relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30."""

    def __init__(self, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
                 max_iteration: int = 30):
        self.relative_fitness = relative_fitness
        self.relative_rmse = relative_rmse
        self.max_iteration = max_iteration


class RegistrationResult:
    def __init__(self, res: "api.IcpResult"):
        self._res = res
        self.transformation = res.transformation
        self.fitness = res.fitness
        self.inlier_rmse = res.inlier_rmse

    @property
    def correspondence_set(self):
        return self._res.correspondence_set

    def __repr__(self):
        return repr(self._res)


def evaluate_registration(source, target, max_correspondence_distance, transformation=None):
    T = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
    if len(source) == 0 or len(target) == 0:
        return RegistrationResult(api.IcpResult(T.copy(), 0.0, 0.0, 0, 0, True))
    return RegistrationResult(api.evaluate_registration(
        source._for_kernel(), target._for_kernel(), max_correspondence_distance, T))


def registration_icp(source, target, max_correspondence_distance, init=None,
                     estimation_method=None, criteria=None):
    if estimation_method is not None and not isinstance(estimation_method,
                                                        TransformationEstimationPointToPoint):
        raise NotImplementedError("only TransformationEstimationPointToPoint is supported")
    c = criteria if criteria is not None else ICPConvergenceCriteria()
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64)
    if len(source) == 0 or len(target) == 0:
        return RegistrationResult(api.IcpResult(T.copy(), 0.0, 0.0, 0, 0, True))
    return RegistrationResult(api.icp(
        source._for_kernel(), target._for_kernel(), T, max_correspondence_distance,
        c.max_iteration, c.relative_fitness, c.relative_rmse))
