"""B200-native batched NCC photo-consistency scorer behind the reference's MVS2 interface."""
