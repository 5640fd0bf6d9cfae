"""pytorch3d.loss.chamfer_distance for the call shape the reference uses (utils.py:207,212)."""
import ptk_b200


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean"):
    if x_lengths is not None or y_lengths is not None or x_normals is not None or y_normals is not None:
        raise NotImplementedError("ptk_b200 chamfer_distance supports homogeneous clouds without normals "
                                  "(the only form pterotactyl calls)")
    if point_reduction != "mean":
        raise NotImplementedError('point_reduction must be "mean"')
    if batch_reduction not in (None, "mean", "sum"):
        raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
    cham, _, _ = ptk_b200.ops.chamfer(x, y)
    if weights is not None:
        cham = cham * weights
    if batch_reduction == "sum":
        cham = cham.sum()
    elif batch_reduction == "mean":
        cham = cham.sum() / (weights.sum() if weights is not None else max(x.shape[0], 1))
    return cham, None
