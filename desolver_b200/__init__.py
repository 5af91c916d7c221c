"""desolver_b200 -- B200-native ensemble engine behind the desolver Python API."""
