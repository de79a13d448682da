"""Drop-in counterparts of bpc.utils (reference: bpc/utils/)."""
