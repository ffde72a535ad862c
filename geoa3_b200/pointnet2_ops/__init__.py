"""Drop-in for the reference's pointnet2_ops package (Model/pointnet2_ops_lib/pointnet2_ops):
`pointnet2_utils` (autograd Functions + grouping modules), `pointnet2_modules` (SA modules) and
`_ext` (the 9 native ops of _ext-src/src/bindings.cpp:6-19), all backed by libgeoa3_b200.so."""
__version__ = "3.0.0+b200"
