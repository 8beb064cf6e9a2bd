"""CPU oracle = test infrastructure (see each module's header). Never imported by yolo_tf_b200."""
