/* openmm_shim: the plugin headers include Lepton but never use it */
