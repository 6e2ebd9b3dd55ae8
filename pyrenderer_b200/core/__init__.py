"""Scene / camera / BSDF objects and the render entry points (reference: core/)."""
