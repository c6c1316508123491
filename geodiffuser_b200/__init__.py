"""geodiffuser_b200 -- B200-native (sm_100a) implementation of GeoDiffuser's geometry-warped
shared-attention hot path behind the reference's attention-controller API.  See DESIGN.md."""
__version__ = "0.1.0"
