"""Host side of the backward kernels (SURVEY 8a').  Filled in by the backward milestone."""


def hs_surface_backward(ctx, grad_out):
    raise NotImplementedError("tg-pose_b200: HSlayer_surface backward kernels are not built yet")


def hs_layer_backward(ctx, grad_out):
    raise NotImplementedError("tg-pose_b200: HS_layer backward kernels are not built yet")


def pool_backward(ctx, g_pooled):
    raise NotImplementedError("tg-pose_b200: Pool_layer backward kernel is not built yet")
