"""Embedder selector with the reference's signature and attributes
(model/custom_embedder_decoder.py:13-164): `Custom_Embedding_Network(input_dims, network_dims,
embed_type, multires, log2_max_hash_size, max_points_per_entry, base_resolution,
desired_resolution, bound)` -> `.forward(x, compute_grad=False)`, `.embeddings_dim`, `.embedder_obj`.

Supported embed_type values: HashGrid, FFB, StyleModNFFB, NerfPos, FourierFeatures, and the two entries the
reference backs with tiny-cuda-nn, HashGridTcnn and FFBTcnn (custom_embedder_decoder.py:153-154) - served here by the
hash-encode kernel's tcnn-semantics mode, without the library (model/embeddings/tcnn_src/).  Anything else raises
ValueError like the reference (:156-157).
"""
import torch
import torch.nn as nn

from .embeddings.frequency_enc import FourierFeature, PositionalEncoding
from .embeddings.hashGridEmbedding import MultiResHashGridMLP


def _nffb_kwargs(input_dims, network_dims, multires, log2_max_hash_size, max_points_per_entry, base_resolution,
                 desired_resolution, bound, style):
    cfg = {'include_input': True, 'in_dim': input_dims, 'network_dims': network_dims, 'n_levels': multires,
           'max_points_per_level': max_points_per_entry, 'log2_hashmap_size': log2_max_hash_size,
           'base_resolution': base_resolution, 'desired_resolution': desired_resolution}
    kw = {'GridEncoderNetConfig': cfg, 'freq_enc_type': 'PositionalEncodingNET', 'has_out': False, 'bound': bound,
          'layers_type': 'SIREN'}
    if style:
        kw['style_modulation'] = True
    return kw


class Custom_Embedding_Network(nn.Module):
    def __init__(self, input_dims, network_dims, embed_type, multires, log2_max_hash_size, max_points_per_entry,
                 base_resolution, desired_resolution, bound):
        super().__init__()
        if embed_type == 'HashGrid':
            obj = MultiResHashGridMLP(True, input_dims, multires, max_points_per_entry, log2_max_hash_size,
                                      base_resolution, desired_resolution)
        elif embed_type in ('FFB', 'StyleModNFFB'):
            from .embeddings.nffb3d import FourierFilterBanks
            obj = FourierFilterBanks(**_nffb_kwargs(input_dims, network_dims, multires, log2_max_hash_size,
                                                    max_points_per_entry, base_resolution, desired_resolution, bound,
                                                    embed_type == 'StyleModNFFB'))
        elif embed_type == 'HashGridTcnn':
            # kwargs of the reference's 'hashGridEncoderTcnn' entry (custom_embedder_decoder.py:83-97)
            from .embeddings.tcnn_src.hashGridEncoderTcnn import MultiResHashGridEncoderTcnn
            obj = MultiResHashGridEncoderTcnn(include_input=True, in_dim=input_dims, network_dims=network_dims,
                                              embed_type='HashGridTcnn', n_levels=multires,
                                              max_points_per_level=max_points_per_entry,
                                              log2_hashmap_size=log2_max_hash_size, base_resolution=base_resolution,
                                              desired_resolution=desired_resolution, grid_embedding_std=0.0001,
                                              per_level_scale=2.0, base_sigma=8.0, exp_sigma=1.26)
        elif embed_type == 'FFBTcnn':
            # kwargs of the reference's 'FFB_TCNN' entry (:41-61): style modulation on
            from .embeddings.tcnn_src.FFB_encoder import FFBEncoder
            kw = _nffb_kwargs(input_dims, network_dims, multires, log2_max_hash_size, max_points_per_entry,
                              base_resolution, desired_resolution, bound, True)
            kw['GridEncoderNetConfig'].update({'embed_type': 'HashGridTcnn', 'base_sigma': 8.0, 'exp_sigma': 1.26,
                                               'grid_embedding_std': 0.0001, 'per_level_scale': 2.0})
            obj = FFBEncoder(**kw)
        elif embed_type == 'NerfPos':
            # the reference passes log2_max_hash_size as max_freq_log2 (custom_embedder_decoder.py:74-81)
            obj = PositionalEncoding(include_input=True, input_dims=input_dims, max_freq_log2=log2_max_hash_size,
                                     num_freqs=multires, log_sampling=True, periodic_fns=[torch.sin, torch.cos])
        elif embed_type == 'FourierFeatures':
            obj = FourierFeature(num_channels=network_dims[0], sigma=1.0, input_dims=input_dims, include_input=True)
        else:
            raise ValueError("Not a valid embedding model type")
        self.embed_type = embed_type
        self.embedder_obj = obj
        self.embeddings_dim = obj.embeddings_dim

    def forward(self, x, compute_grad=False):
        return self.embedder_obj.forward(x)
