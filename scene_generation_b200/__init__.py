"""B200-native hot path of ashual/scene_generation."""
__version__ = '0.1.0'
