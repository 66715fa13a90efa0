"""B200-native RSSM latent-dynamics hot path (sm_100a CUDA behind a C ABI)."""
__all__ = ["engine", "_lib"]
