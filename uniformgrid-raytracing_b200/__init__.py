"""B200-native per-frame hot path of sushruta/uniformgrid-raytracing."""
