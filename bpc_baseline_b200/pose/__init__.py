"""Consumer side of the hot path: the pose head that reads the crop tensor."""
