"""Decoder structure dicts used by the reference's evaluation scripts (config literals, not code)."""

# test_modelnet_VAE_dr.py:176-185 (latent dim 64)
MODELNET_DECODER = {
    'name': 'decoder',
    'input_dim': 64,
    'output_shape': [64, 64, 64, 1],
    'filter_num_list': [512, 256, 128, 64, 1],
    'filter_size_list': [4, 4, 4, 4, 4],
    'strides_list': [1, 2, 2, 2, 2],
    'activation': 'elu',
    'final_activation': 'sigmoid',
}
# test_pascal_VAE_dr.py:196-205 (latent dim 16)
PASCAL_DECODER = dict(MODELNET_DECODER, input_dim=16)
# test_modelnet_VAE_dr.py:172-181 (voxel encoder of the ModelNet VAE, latent dim 64 -> 2 * 64 output channels)
MODELNET_ENCODER = {
    'name': 'encoder3D',
    'input_shape': [64, 64, 64, 1],
    'filter_num_list': [64, 128, 256, 512, 128],
    'filter_size_list': [4, 4, 4, 4, 4],
    'strides_list': [2, 2, 2, 2, 1],
    'final_pool': 'average',
    'activation': 'elu',
    'final_activation': 'None',
}
# test_pascal_VAE_dr.py:186-195 (encoder backbone / head of the Pascal3D VAE; head built with last_pooling='max',
# src/module/nolbo.py:778-783)
PASCAL_ENCODER_BACKBONE = {'name': 'nolbo_backbone', 'z_dim': 16, 'activation': 'elu'}
PASCAL_ENCODER_HEAD = {'name': 'nolbo_head', 'output_dim': 32, 'filter_num_list': [], 'filter_size_list': [],
                       'activation': 'elu'}
PASCAL_IMAGE_SIZE = (256, 256)   # test_pascal_VAE_dr.py:52

FLOP_PER_IMAGE_ENCODER = 7.162e9   # 2 x 3.581 GMAC at 256 x 256, interior + border taps
FLOP_PER_DECODE = {64: 6.663830528e9, 16: 6.663781376e9}  # exact MAC*2, SURVEY.md section 7
