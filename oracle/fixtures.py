"""Deterministic test cases shared by the golden generator, the oracle tests and the
GPU parity tests -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Nothing here reads /root/reference.  Parameters and inputs are drawn from
``numpy.random.RandomState`` (frozen legacy stream) so that the generator (which runs
the real reference in the build container) and the tests (which run anywhere)
construct bit-identical state_dicts without shipping them.
"""
import numpy as np

# name -> dict(model kwargs, input shape, mode, loss)
CASES = {
    # BASELINE cfg 1: the reference's own CPU-runnable case
    'cfg1_eval': dict(model=dict(in_channels=1, out_channels=2, dim=3, n_blocks=2, start_filts=8),
                      x=(1, 1, 32, 32, 32), train=False),
    'cfg1_train': dict(model=dict(in_channels=1, out_channels=2, dim=3, n_blocks=2, start_filts=8),
                       x=(2, 1, 16, 16, 16), train=True),
    # BASELINE cfg 2 (GroupNorm, 3 blocks) at reduced width / extent
    'cfg2_sf8_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group'),
                           x=(2, 1, 16, 16, 16), train=True),
    # BASELINE cfg 2 at its real width (32/64/128 channels: the tensor-core shapes)
    'cfg2_sf32_train': dict(model=dict(n_blocks=3, start_filts=32, normalization='group'),
                            x=(1, 1, 16, 16, 16), train=True),
    'cfg2_sf32_eval24': dict(model=dict(n_blocks=3, start_filts=32, normalization='group'),
                             x=(1, 1, 24, 24, 24), train=False),
    # BASELINE cfg 3: anisotropic planar blocks, BN in train mode
    'cfg3_planar_train': dict(model=dict(n_blocks=3, start_filts=8, planar_blocks=(0, 1)),
                              x=(1, 1, 4, 32, 32), train=True),
    # BASELINE cfg 5: 2D path (D = 1)
    'cfg5_2d_train': dict(model=dict(dim=2, n_blocks=3, start_filts=8),
                          x=(2, 1, 32, 32), train=True),
    # odd extents: ceil_mode pooling + autocrop (unet.py:256-325)
    'odd_eval': dict(model=dict(n_blocks=3, start_filts=8, normalization='group'),
                     x=(1, 1, 11, 13, 18), train=False),
    # multi-channel input / more classes / no norm
    'c3_none_train': dict(model=dict(in_channels=3, out_channels=4, n_blocks=2, start_filts=8,
                                     normalization='none'),
                          x=(1, 3, 8, 16, 16), train=True),
    # VALID convolutions (oracle pin only)
    'valid_eval': dict(model=dict(n_blocks=2, start_filts=8, normalization='group', conv_mode='valid'),
                       x=(1, 1, 20, 20, 20), train=False),
}

# Options off the default configuration: goldens from the real reference pin the plain-torch
# restatement (oracle/torch_ref.py) and the CUDA path; the C/numpy oracle covers the default configuration only.
OPTION_CASES = {
    'opt_leaky_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group', activation='leaky'),
                            x=(2, 1, 16, 16, 16), train=True),
    'opt_prelu_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group', activation='prelu'),
                            x=(2, 1, 16, 16, 16), train=True),
    'opt_prelu_bn_eval': dict(model=dict(n_blocks=2, start_filts=8, activation='prelu'),
                              x=(1, 1, 16, 16, 16), train=False),
    'opt_silu_train': dict(model=dict(n_blocks=2, start_filts=8, activation='silu'),
                           x=(2, 1, 16, 16, 16), train=True),
    'opt_lin_none_train': dict(model=dict(n_blocks=2, start_filts=8, normalization='none', activation='lin'),
                               x=(1, 1, 8, 16, 16), train=True),
    'opt_silu_eval': dict(model=dict(n_blocks=2, start_filts=8, activation='silu'),
                          x=(1, 1, 16, 16, 16), train=False),
    'opt_rrelu_eval': dict(model=dict(n_blocks=2, start_filts=8, normalization='group', activation='rrelu'),
                           x=(1, 1, 16, 16, 16), train=False),
    'opt_add_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group', merge_mode='add'),
                          x=(2, 1, 16, 16, 16), train=True),
    # (28^3 in -> 12^3 out: with a 4^3 output one ReLU whose pre-activation sits at 0 +- TF32 noise moves a bias gradient by 1/64)
    'opt_add_valid_train': dict(model=dict(n_blocks=2, start_filts=8, normalization='group', merge_mode='add',
                                           conv_mode='valid'),
                                x=(1, 1, 28, 28, 28), train=True),
    'opt_resize_nearest_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group',
                                                up_mode='resizeconv_nearest'),
                                     x=(2, 1, 16, 16, 16), train=True),
    'opt_resize_linear_train': dict(model=dict(n_blocks=2, start_filts=8, up_mode='resizeconv_linear'),
                                    x=(2, 1, 8, 16, 16), train=True),
    'opt_resize_nearest1_planar_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group',
                                                        up_mode='resizeconv_nearest1', planar_blocks=(0,)),
                                             x=(1, 1, 8, 16, 16), train=True),
    'opt_resize_linear1_2d_train': dict(model=dict(dim=2, n_blocks=2, start_filts=8, up_mode='resizeconv_linear1'),
                                        x=(2, 1, 16, 16), train=True),
    # odd extents: ceil-mode pooling makes the up-sampled tensor one voxel too large -> autocrop of the resize-conv output
    'opt_resize_linear_odd_train': dict(model=dict(n_blocks=3, start_filts=8, normalization='group',
                                                   up_mode='resizeconv_linear'),
                                        x=(1, 1, 11, 13, 18), train=True),
}

# elektronn3.models.resunet.UNet: `arch='resunet'` selects the model class
RESUNET_CASES = {
    # no residual blocks: the arithmetic of the plain UNet under resunet's parameter names
    'res0_train': dict(arch='resunet', model=dict(n_blocks=3, start_filts=8, normalization='group'),
                       x=(2, 1, 16, 16, 16), train=True),
    # one residual ConvBlock per level: projection shortcuts (channel counts differ everywhere but at the image)
    'res1_train': dict(arch='resunet', model=dict(n_blocks=3, start_filts=8, normalization='group', enc_res_blocks=1,
                                                  dec_res_blocks=1),
                       x=(2, 1, 16, 16, 16), train=True),
    # two per level: identity shortcuts as well, BatchNorm in training mode, a planar block
    'res2_bn_train': dict(arch='resunet', model=dict(n_blocks=2, start_filts=8, enc_res_blocks=2, dec_res_blocks=2,
                                                     planar_blocks=(0,)),
                          x=(2, 1, 8, 16, 16), train=True),
    # eval-mode BatchNorm around the shortcuts (never folded into the weights), odd extents
    'res2_bn_eval': dict(arch='resunet', model=dict(n_blocks=3, start_filts=8, enc_res_blocks=2, dec_res_blocks=1),
                         x=(1, 1, 11, 13, 18), train=False),
    # merge_mode='add' (identity shortcut on the sum), leaky, resize-conv up-sampling in a separate case
    'res1_add_leaky_train': dict(arch='resunet', model=dict(n_blocks=2, start_filts=8, normalization='group', enc_res_blocks=1,
                                                            dec_res_blocks=1, merge_mode='add', activation='leaky'),
                                 x=(1, 1, 16, 16, 16), train=True),
    'res1_resize_none_train': dict(arch='resunet', model=dict(n_blocks=2, start_filts=8, normalization='none',
                                                              enc_res_blocks=1, dec_res_blocks=2,
                                                              up_mode='resizeconv_linear'),
                                   x=(1, 1, 8, 16, 16), train=True),
}

# Predictor / tiled_apply cases: model, volume shape, tile, overlap
PRED_CASES = {
    'pred_small': dict(model=dict(n_blocks=2, start_filts=8), vol=(1, 1, 16, 24, 16),
                       tile=(8, 8, 8), overlap=(4, 4, 4), out_channels=2),
    'pred_sf32': dict(model=dict(n_blocks=2, start_filts=32), vol=(1, 1, 16, 16, 32),
                      tile=(8, 8, 16), overlap=(4, 4, 8), out_channels=2),
}


def state_shapes_from_torch(model):
    """helper for the generator only: ordered (key, shape, dtype-kind) of a torch module"""
    return [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]


def make_state(shapes, seed=1234):
    """Deterministic state_dict (numpy) for a list of (key, shape, dtype) entries.

    conv / convT weights ~ N(0, xavier std) as in UNet.weight_init (unet.py:885-892)
    but biases and norm affine parameters are made NON-trivial (the reference inits
    them to 0 / 1, which would hide bias / gamma / beta bugs); running stats are
    non-trivial too so that eval-mode BN folding is exercised.
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape, dt in shapes:
        if key.endswith('num_batches_tracked'):
            sd[key] = np.array(3, dtype=np.int64)
        elif key.endswith('running_mean'):
            sd[key] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
        elif key.endswith('running_var'):
            sd[key] = (0.5 + rs.random_sample(shape)).astype(np.float32)
        elif '.act' in key and key.endswith('.weight'):      # nn.PReLU slope (get_activation 'prelu', unet.py:189-190)
            sd[key] = (0.25 + 0.1 * rs.standard_normal(shape)).astype(np.float32)
        elif key.endswith('.weight') and len(shape) >= 3:
            rf = int(np.prod(shape[2:]))
            std = np.sqrt(2.0 / ((shape[0] + shape[1]) * rf))
            sd[key] = (std * rs.standard_normal(shape)).astype(np.float32)
        elif key.endswith('.weight'):      # norm gamma
            sd[key] = (1.0 + 0.2 * rs.standard_normal(shape)).astype(np.float32)
        else:                              # conv bias / norm beta
            sd[key] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
    return sd


def make_input(shape, seed=77, kind='randn'):
    rs = np.random.RandomState(seed)
    if kind == 'randn':
        return rs.standard_normal(shape).astype(np.float32)
    # neuro_data_cdhw-like: smooth uint8 texture normalised with the reference's
    # mean/std (examples/train_unet_neurodata.py:150-151)
    raw = rs.standard_normal(shape)
    for ax in range(2, len(shape)):
        raw = raw + np.roll(raw, 1, axis=ax) + np.roll(raw, -1, axis=ax)
    raw = raw / raw.std()
    u8 = np.clip(155.291411 + 42.599973 * raw, 0, 255).astype(np.uint8)
    return ((u8.astype(np.float32) - 155.291411) / 42.599973).astype(np.float32)


def make_target(shape_x, n_classes, seed=99):
    rs = np.random.RandomState(seed)
    return rs.randint(0, n_classes, size=(shape_x[0],) + tuple(shape_x[2:])).astype(np.int64)


def digest(a, n=512):
    """Small fingerprint of a big array: strided sample + norms (pins parity without
    committing megabytes)."""
    f = np.asarray(a, dtype=np.float32).ravel()
    stride = max(1, f.size // n)
    return np.concatenate([f[::stride][:n].astype(np.float64),
                           [f.astype(np.float64).sum(), np.abs(f.astype(np.float64)).sum(),
                            np.sqrt((f.astype(np.float64) ** 2).sum())]])
