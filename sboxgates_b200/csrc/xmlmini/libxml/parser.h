/* xmlmini/libxml/parser.h -- the part of libxml2's parser.h state.c needs.  See tree.h. */
#ifndef SBG_STUB_LIBXML_PARSER_H
#define SBG_STUB_LIBXML_PARSER_H
#include "tree.h"
xmlDocPtr xmlParseFile(const char *filename);
#endif
