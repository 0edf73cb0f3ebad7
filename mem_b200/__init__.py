"""mem_b200 -- B200-native (sm_100a) implementation of the MEM pretraining hot path.

Drop-in for the one data-parallel hot path of tum-vision/mem: event rasterisation
(``process_data.histogram`` / ``datasets.EventArrToImg``), dVAE visual tokens
(``vae_model.DiscreteVAE.get_codebook_indices``), blockwise masks
(``masking_generator.MaskingGenerator``), the masked ViT
(``modeling_pretrain``: ``pt_vit`` / ``beit_*_8k_vocab``) and
``engine_for_pretraining.train_one_epoch``.  Python/PyTorch is the host layer;
all device work goes through the C ABI of ``libmemb.so`` (``include/memb.h``).
"""
__version__ = "0.1.0"
