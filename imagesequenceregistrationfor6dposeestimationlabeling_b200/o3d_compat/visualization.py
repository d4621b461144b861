"""o3d.visualization subset: the reference opens blocking GUI windows; no-ops here."""


def draw_geometries(geometries, *args, **kwargs):
    return None
