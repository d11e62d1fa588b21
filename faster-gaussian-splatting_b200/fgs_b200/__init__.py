"""fgs_b200: Blackwell-native (sm_100a) drop-in for the FasterGSCudaBackend operator API."""
