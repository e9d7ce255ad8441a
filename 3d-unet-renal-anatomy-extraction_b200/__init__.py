"""B200-native 3D U-Net hot path."""
