"""Scene input (reference: io_utils/)."""
