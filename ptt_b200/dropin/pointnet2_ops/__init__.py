"""Drop-in stand-in for the third-party `pointnet2_ops` package (erikwijmans/Pointnet2_PyTorch) that
the reference imports at ptt/models/backbones_3d/pointnet2/pointnet2_utils.py:24.  Only `_ext` -- the
native-op module the reference actually uses -- is provided; it is backed by libptt_b200.so."""
__version__ = "3.0.0+ptt_b200"
