"""Reference-side bindings: what a maintainer of the reference adds to run its Python on libmode_b200."""
