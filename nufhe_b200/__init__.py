"""nufhe_b200 -- B200-native engine for the gate-bootstrapping hot path of nucypher/nufhe."""
