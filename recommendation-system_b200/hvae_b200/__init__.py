"""hvae_b200 -- B200-native HybridVAE training / evaluation hot path."""
