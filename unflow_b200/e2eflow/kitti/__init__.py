"""KITTI data directories and input queues (reference src/e2eflow/kitti/)."""
