"""TEST INFRASTRUCTURE: stand-in for the `trimesh` package (absent from this image, no network).  The reference only
needs the name `trimesh.base.Trimesh` to exist (type annotations / optional constructors)."""
