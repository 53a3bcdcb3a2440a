"""The reference's python binding surface for the conv library, on libconvnet_b200.so.

Same function names and argument meaning as cudamat/cudamat_conv_gemm.py:61-157
(`convUp(images, filters, targets, conv_desc, scaleTargets=0)` ...), so the
parity tests read like py/test_conv.py.  Arguments are CUDAMatrix objects.
"""
from . import lib as _lib
from .abi import ConvDesc


def _L():
    return _lib.load()


def convUp(images, filters, targets, conv_desc, scaleTargets=0):
    _L().convUpGemm(images.p_mat, filters.p_mat, targets.p_mat, images.p_shape4d, filters.p_shape4d,
                    targets.p_shape4d, conv_desc, scaleTargets)


def convDown(hidSums, filters, targets, conv_desc, scaleTargets=0):
    _L().convDownGemm(hidSums.p_mat, filters.p_mat, targets.p_mat, hidSums.p_shape4d, filters.p_shape4d,
                      targets.p_shape4d, conv_desc, scaleTargets)


def convOutp(images, hidSums, targets, conv_desc, scaleTargets=0, scaleGradients=1):
    _L().convOutpGemm(images.p_mat, hidSums.p_mat, targets.p_mat, images.p_shape4d, hidSums.p_shape4d,
                      targets.p_shape4d, conv_desc, scaleTargets, scaleGradients)


def localUp(images, filters, targets, conv_desc, scaleTargets=0):
    _L().localUpGemm(images.p_mat, filters.p_mat, targets.p_mat, images.p_shape4d, filters.p_shape4d,
                     targets.p_shape4d, conv_desc, scaleTargets)


def localDown(hidSums, filters, targets, conv_desc, scaleTargets=0):
    _L().localDownGemm(hidSums.p_mat, filters.p_mat, targets.p_mat, hidSums.p_shape4d, filters.p_shape4d,
                       targets.p_shape4d, conv_desc, scaleTargets)


def localOutp(images, hidSums, targets, conv_desc, scaleTargets=0, scaleGradients=1):
    _L().localOutpGemm(images.p_mat, hidSums.p_mat, targets.p_mat, images.p_shape4d, hidSums.p_shape4d,
                       targets.p_shape4d, conv_desc, scaleTargets, scaleGradients)


def MaxPool(images, targets, conv_desc):
    _L().MaxPoolGemm(images.p_mat, targets.p_mat, images.p_shape4d, targets.p_shape4d, conv_desc, 0.0, 1.0)


def MaxPoolUndo(images, grad, maxes, targets, conv_desc, scaleTargets=0):
    _L().MaxPoolUndoGemm(images.p_mat, grad.p_mat, maxes.p_mat, targets.p_mat, images.p_shape4d, grad.p_shape4d,
                         conv_desc, scaleTargets)


def AvgPool(images, targets, conv_desc):
    _L().AvgPoolGemm(images.p_mat, targets.p_mat, images.p_shape4d, targets.p_shape4d, conv_desc, 0.0, 1.0)


def AvgPoolUndo(avgGrads, targets, conv_desc, scaleTargets=0):
    _L().AvgPoolUndoGemm(avgGrads.p_mat, targets.p_mat, avgGrads.p_shape4d, targets.p_shape4d, conv_desc,
                         scaleTargets)


def UpSample(images, targets, factor, scaleTargets=0):
    _L().UpSampleGemm(images.p_mat, targets.p_mat, images.p_shape4d, targets.p_shape4d, factor, scaleTargets)


def DownSample(images, targets, factor):
    _L().DownSampleGemm(images.p_mat, targets.p_mat, images.p_shape4d, targets.p_shape4d, factor)


def ResponseNormCrossMap(images, targets, sizeF, addScale, powScale, blocked):
    _, _, _, num_filters = images.shape4d
    _L().ResponseNormCrossMapGemm(images.p_mat, targets.p_mat, num_filters, sizeF, addScale, powScale, blocked)


def ResponseNormCrossMapUndo(derivs, images, targets, sizeF, addScale, powScale, blocked):
    _, _, _, num_filters = images.shape4d
    _L().ResponseNormCrossMapUndoGemm(derivs.p_mat, images.p_mat, targets.p_mat, num_filters, sizeF, addScale,
                                      powScale, blocked)


def ResponseNormCrossMap3D(images, targets, sizeF, addScale, powScale, blocked, image_size_t):
    num_filters = images.shape4d[3] // image_size_t
    _L().ResponseNormCrossMap3DGemm(images.p_mat, targets.p_mat, num_filters, sizeF, addScale, powScale, blocked,
                                    image_size_t)


def ResponseNormCrossMap3DUndo(derivs, images, targets, sizeF, addScale, powScale, blocked, image_size_t):
    num_filters = images.shape4d[3] // image_size_t
    _L().ResponseNormCrossMap3DUndoGemm(derivs.p_mat, images.p_mat, targets.p_mat, num_filters, sizeF, addScale,
                                        powScale, blocked, image_size_t)


def convUp3D(images, filters, targets, conv_desc, scaleTargets=0):
    _L().convUp3DGemm(images.p_mat, filters.p_mat, targets.p_mat, images.p_shape4d, filters.p_shape4d,
                      targets.p_shape4d, conv_desc, scaleTargets)


def convDown3D(hidSums, filters, targets, conv_desc, scaleTargets=0):
    _L().convDown3DGemm(hidSums.p_mat, filters.p_mat, targets.p_mat, hidSums.p_shape4d, filters.p_shape4d,
                        targets.p_shape4d, conv_desc, scaleTargets)


def convOutp3D(images, hidSums, targets, conv_desc, scaleTargets=0, scaleGradients=1):
    _L().convOutp3DGemm(images.p_mat, hidSums.p_mat, targets.p_mat, images.p_shape4d, hidSums.p_shape4d,
                        targets.p_shape4d, conv_desc, scaleTargets, scaleGradients)


# 3-D pooling goes through the same entry points (kernel_size_t/stride_t in the descriptor)
MaxPool3D, MaxPool3DUndo, AvgPool3D, AvgPool3DUndo = MaxPool, MaxPoolUndo, AvgPool, AvgPoolUndo


# ---- ABI-2 (cudamat/cudamat_conv.py): only the entry points whose signature differs ------------
def convOutpPartial(images, hidSums, targets, conv_desc, partialSumY, partialSumX, scaleTargets=0,
                    scaleGradients=1):
    _L().convOutp(images.p_mat, hidSums.p_mat, targets.p_mat, images.p_shape4d, hidSums.p_shape4d,
                  targets.p_shape4d, conv_desc, partialSumY, partialSumX, scaleTargets, scaleGradients)
