"""B200-native RSSM latent-dynamics hot path."""
__version__ = "0.2.0"


def install(agent, **kw):
    """Swap the RSSM hot path of a reference `Dreamer` for this library (safe_dreamer_b200/installer.py)."""
    from .installer import install as _install
    return _install(agent, **kw)
