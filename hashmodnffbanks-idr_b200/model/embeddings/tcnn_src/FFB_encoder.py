"""`FFBEncoder` with the reference's constructor (model/embeddings/tcnn_src/FFB_encoder.py:21-203), on the
tiny-cuda-nn-free grid of hashGridEncoderTcnn.py.  Same filter-bank data flow as nffb3d.FourierFilterBanks with the
three differences the reference file has: the grid is the tcnn-semantics encoder, chunk i is level i's F features
(`grid_x.view(-1, L, F)`, :139-141), and the trunk width is the positional encoding's width itself (:75-78), not twice it.
"""
from ..nffb3d import FourierFilterBanks
from .hashGridEncoderTcnn import MultiResHashGridEncoderTcnn


class FFBEncoder(FourierFilterBanks):
    WIDTH_MULT = 1

    def _make_grid(self, cfg):
        return MultiResHashGridEncoderTcnn(
            include_input=cfg['include_input'], in_dim=cfg['in_dim'], network_dims=cfg.get('network_dims'),
            embed_type='HashGridTcnn', n_levels=cfg['n_levels'], max_points_per_level=cfg['max_points_per_level'],
            log2_hashmap_size=cfg['log2_hashmap_size'], base_resolution=cfg['base_resolution'],
            desired_resolution=cfg['desired_resolution'], base_sigma=cfg.get('base_sigma', 8.0),
            exp_sigma=cfg.get('exp_sigma', 1.26), grid_embedding_std=cfg.get('grid_embedding_std', 1e-4),
            per_level_scale=cfg.get('per_level_scale', 2.0))

    def _chunk_width(self):
        return self.max_points_per_level
