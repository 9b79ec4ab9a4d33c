"""flipped-vqa-b200: the LLaMA-VQA training step of inesriahi/Flipped-VQA on hand-written sm_100a kernels."""
__version__ = "0.1.0"
