from quadtree_mpnnlstm_b200.graph_functions import (CONDITIONS, Graph, Mesh, create_static_heterogeneous_graph,  # noqa: F401
                                                    create_static_homogeneous_graph, flatten, image_to_graph,
                                                    image_to_graph_pixelwise, plot_contours, unflatten)
