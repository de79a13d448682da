"""Drop-in counterparts of bpc.inference (reference: bpc/inference/)."""
