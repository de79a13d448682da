"""Drop-in counterparts of bpc.inference.utils (reference: bpc/inference/utils/)."""
