/* oracle/stubs/libxml/parser.h -- TEST INFRASTRUCTURE, not product code.  See tree.h. */
#ifndef SBG_STUB_LIBXML_PARSER_H
#define SBG_STUB_LIBXML_PARSER_H
#include "tree.h"
xmlDocPtr xmlParseFile(const char *filename);
#endif
