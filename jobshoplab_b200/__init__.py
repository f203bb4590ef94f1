"""jobshoplab_b200: B200-native batched JobShopLab env-step path."""
