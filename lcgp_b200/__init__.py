"""lcgp_b200 -- B200-native (sm_100a) implementation of LCGP's emulator-fitting hot path."""
