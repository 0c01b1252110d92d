"""devias_b200 -- B200-native (sm_100a) implementation of the DEVIAS hot path behind the reference API."""
__version__ = '0.1.0'
