"""Hyper-parameters of the SPAIR drop-in, under the reference's constant names.

Mirror of the reference's ``spair/config.py:1-81`` (module-level constants, no parsing, no
overrides): the same names carry the same values, because every shape in the hot path is
derived from them and ``train.py`` reads ``cfg.INPUT_IMAGE_SHAPE`` / ``cfg.BATCH_SIZE``.
Unlike the reference, the model reads these at CONSTRUCTION time (not as import-time default
arguments), so ``cfg.X = ...`` followed by ``SPAIR(...)`` is enough to build another shape.
"""
import os

# --- data / batch -------------------------------------------------------------------------
BATCH_SIZE = 32
INPUT_IMAGE_SHAPE = [1, 128, 128]            # [C, H, W]


def _conv(filters, kernel_size, stride):
    return dict(filters=filters, kernel_size=kernel_size, stride=stride)


# --- network topologies -------------------------------------------------------------------
DEFAULT_MLP_TOPOLOGY = [100, 100]
DEFAULT_BACKBONE_TOPOLOGY = [_conv(128, 4, 3), _conv(128, 4, 2), _conv(128, 4, 2),
                             _conv(128, 1, 1), _conv(128, 1, 1), _conv(128, 1, 1)]
# stride-2/2/2 variant: 8-px cells (BASELINE.json configs 3 and 4)
CELL8_BACKBONE_TOPOLOGY = [_conv(128, 4, 2), _conv(128, 4, 2), _conv(128, 4, 2),
                           _conv(128, 1, 1), _conv(128, 1, 1), _conv(128, 1, 1)]
# kept for name compatibility; the conv object encoder/decoder of the reference is dead code
CONV_OBJECT_ENCODER_TOPOLOGY = [_conv(32, 4, 2), _conv(32, 3, 2), _conv(32, 3, 2), _conv(32, 1, 1)]

N_BACKBONE_FEATURES = 100
N_PASSTHROUGH_FEATURES = 100

# --- latent sizes -------------------------------------------------------------------------
N_ATTRIBUTES = 50
N_CONTEXT_DIM = 4 + N_ATTRIBUTES + 1 + 1     # box, attr, depth, pres
N_LOOKBACK = 1                                # neighbourhood radius of the lateral context

OBJECT_SHAPE = [28, 28]                       # glimpse size
ANCHORBOX_SHAPE = [48, 48]

# box range relative to the cell / anchor
MAX_YX = 1.5
MIN_YX = -0.5
MAX_HW = 1.0
MIN_HW = 0.0

# --- priors {name: [mean, std]} -------------------------------------------------------------
PRIORS = {
    'cy_logit': [0., 1.],
    'cx_logit': [0., 1.],
    'height_logit': [7.00, 0.5],
    'width_logit': [7.00, 0.5],
    'attr': [0., 1.],
    'depth_logit': [0., 1.],
}
VAE_BETA = 1

# --- schedules ----------------------------------------------------------------------------
LATENT_VAR_TRAINING_WHEEL_PARAM = dict(start=1.0, end=0.0, decay_rate=0.0, decay_step=1000., staircase=True)
OBJ_PRES_COUNT_LOG_PRIOR = dict(start=1000000.0, end=0.0125, decay_rate=0.1, decay_step=1000., log_space=True)

# --- decoder output scaling ---------------------------------------------------------------
OBJ_LOGIT_SCALE = 2.0
ALPHA_LOGIT_SCALE = 0.1
ALPHA_LOGIT_BIAS = 5.0

# --- environment --------------------------------------------------------------------------
IS_LOCAL = 'LOCAL' in os.environ
# B200 build only: print the per-step loss breakdown like the reference does (each print is a
# device->host sync, so it is off unless asked for)
VERBOSE = 'SPAIR_VERBOSE' in os.environ
