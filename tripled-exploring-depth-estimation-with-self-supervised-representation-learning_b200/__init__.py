"""B200-native fused view-synthesis loss."""
