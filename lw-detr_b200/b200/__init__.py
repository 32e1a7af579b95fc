"""Host-side glue of the B200-native LW-DETR path (ctypes binding, configs, synthetic weights)."""
