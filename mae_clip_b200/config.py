"""Defaults of the drop-in classes.

Same values as the reference's module constants (``/root/reference/config.py:21-32``), which the
reference freezes into constructor defaults at import time (``CLIP.py:12-14``, ``modules.py:59-60``).
Only the constants the hot path reads are kept; paths, epochs and optimiser settings belong to the
training driver, which is out of scope.
"""
model_name = "resnet50"
image_embedding = 2048
text_encoder_model = "distilbert-base-uncased"
text_embedding = 768
max_length = 200

pretrained = True
trainable = True
temperature = 1.0

size = 224

projection_dim = 256
dropout = 0.1

# engine of the contrastive-loss / projection GEMMs: "simt_fp32", "tc_f16x3" or "tc_f16"
gemm_mode = "tc_f16x3"
