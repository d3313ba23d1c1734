"""Model registry for the TAI path: the keys of the reference's ``src/models/create_model.py`` that name
a model on this path (create_model.py:27-36, 69-76) with identical constructor arguments.  Keys of the
out-of-scope model families (SCT*, OFFillInModel, bi-SA, bi-TWA, TW_P_F) raise like an unknown key."""
from .mcnet.mcnet import MCNetFillInModel
from .slomo.slomo import SloMoFillInModel
from .tai.tai import TAIFillInModel
from .twi.twi import TimeWeightedInterpolationFillInModel

_REGISTRY = {
    'TAI_gray': lambda: TAIFillInModel(64, 1, 3, 51, num_block=5),
    'TAI_color': lambda: TAIFillInModel(64, 3, 3, 51, num_block=4),
    'MCNet_gray': lambda: MCNetFillInModel(64, 1, 3),
    'MCNet_color': lambda: MCNetFillInModel(64, 3, 3),
    'SloMoFillInModel_color': lambda: SloMoFillInModel(32, 3),
    'SloMoFillInModel_gray': lambda: SloMoFillInModel(32, 1),
    'TimeWeightedInterpolationFillInModel_gray': lambda: TimeWeightedInterpolationFillInModel(64, 1, 3, 51, num_block=5),
    'TimeWeightedInterpolationFillInModel_color': lambda: TimeWeightedInterpolationFillInModel(64, 3, 3, 51, num_block=4),
}


def create_model(model_key):
    """model_key -> freshly constructed (CPU, un-initialised) model."""
    try:
        return _REGISTRY[model_key]()
    except KeyError:
        raise RuntimeError('model key %r is not on the TAI hot path; known keys: %s'
                           % (model_key, ', '.join(sorted(_REGISTRY))))
