"""Decoder structure dicts used by the reference's evaluation scripts (config literals, not code)."""

# test_modelnet_VAE_dr.py:176-185 (latent dim 64)
MODELNET_DECODER = {
    'name': 'docoder',
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

FLOP_PER_DECODE = {64: 6.663830528e9, 16: 6.663781376e9}  # exact MAC*2, SURVEY.md section 7
