"""Host-side mirror of the reference package ``e2eflow`` (src/e2eflow/) for the hot path:
``ops`` and ``core.{flownet, losses, image_warp, unsupervised, util}`` with the reference's
function names, argument names, defaults, return structures and NHWC layouts, operating on
torch CUDA tensors and backed by libunflow.so."""
