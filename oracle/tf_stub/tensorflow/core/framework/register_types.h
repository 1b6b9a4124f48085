// Stand-in (empty): the reference's CUDA files include it but use nothing from it.
