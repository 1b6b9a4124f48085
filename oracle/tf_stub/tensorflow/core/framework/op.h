// Stand-in (empty): included by correlation_op.h, nothing of it is used by the CUDA files.
